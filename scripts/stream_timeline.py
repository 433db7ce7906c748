import os, sys
os.environ["MBQC_STREAM_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import mentpy_b200 as mb
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
w = int(sys.argv[1]); fuse = int(sys.argv[2])
gs = mb.templates.linear_cluster(w + 16)
ps = mb.PatternSimulator(gs, backend="cuda-sv-stream", window_size=w, fuse=fuse)
ang = np.random.default_rng(4).uniform(0, 2 * np.pi, w + 15)
ps.run(ang); ps.run(ang)
if rank == 0:
    from mentpy_b200.streaming import StreamExecutor
    # the executor object is recreated per run: re-run once more by hand to keep the timeline
    sim = ps.simulator
    ex = StreamExecutor(sim.plan, sim._engine, rank, world.bit_length() - 1, sim.fuse)
else:
    sim = ps.simulator
    ex = __import__("mentpy_b200.streaming", fromlist=["StreamExecutor"]).StreamExecutor(sim.plan, sim._engine, rank, world.bit_length() - 1, sim.fuse)
ex.run(ang)
if rank == 0:
    tot = 0
    for label, ms in ex.timeline:
        tot += ms; print(f"{ms:9.3f} ms  {label}")
    print("total", tot)
if world > 1: dist.destroy_process_group()
