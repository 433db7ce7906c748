"""Run under torchrun (one process per GPU): sharded streaming state-vector vs oracles.
Prints 'STREAM_CHECK_OK' from rank 0 on success."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import mentpy_b200 as mb
from oracle import matrix_free
from oracle.pattern_data import PatternData


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    ok = True
    cases = [("linear_cluster", [22], 12, 4), ("grid_cluster", [3, 7], 9, 3), ("grid_cluster", [2, 12], 11, 5),
             ("linear_cluster", [30], 16, 4)]
    for name, args, w, fuse in cases:
        gs = getattr(mb.templates, name)(*args)
        ang = np.random.default_rng(17).uniform(0, 2 * np.pi, len(gs.trainable_nodes))
        ps = mb.PatternSimulator(gs, backend="cuda-sv-stream", window_size=w, fuse=fuse)
        got = ps.run(ang, output_form="sv")
        want = matrix_free.run_sv_batch(PatternData.from_circuit(gs), ang, window_size=w)[0]
        infid = abs(1 - abs(np.vdot(got, want)) ** 2)
        good = infid < 1e-10 and np.allclose(got, want, atol=1e-9)
        if rank == 0:
            print(f"{name}{args} w={w} fuse={fuse} world={world}: infidelity {infid:.2e} {'ok' if good else 'FAIL'}", flush=True)
        ok &= good
    # large window: analytic oracle only (the reference stops at w ~ 12)
    w = int(os.environ.get("STREAM_BIG_W", "26"))
    gs = mb.templates.linear_cluster(w + 16)
    ang = np.random.default_rng(4).uniform(0, 2 * np.pi, w + 15)
    ps = mb.PatternSimulator(gs, backend="cuda-sv-stream", window_size=w, fuse=4)
    got = ps.run(ang, output_form="sv")
    want = matrix_free.linear_cluster_analytic(ang)[0]
    infid = abs(1 - abs(np.vdot(got, want)) ** 2)
    if rank == 0:
        print(f"linear_cluster({w + 16}) w={w} world={world}: infidelity vs analytic {infid:.2e}", flush=True)
    ok &= infid < 1e-10
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print("STREAM_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
