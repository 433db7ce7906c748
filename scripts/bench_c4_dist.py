"""BASELINE config 4 under torchrun: parameter-shift gradients of grid_cluster(4,5) for 2^20 angle
vectors split contiguously across the ranks (no collective on the data path; max over ranks)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import mentpy_b200 as mb
from mentpy_b200.dist import slice_bounds
from mentpy_b200.gradients import psr_gradient_batched

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
gs = mb.templates.grid_cluster(4, 5); T = len(gs.trainable_nodes)
ps = mb.PatternSimulator(gs, backend="cuda-sv")
B = 1 << 20
lo, hi = slice_bounds(B, rank, world)
gen = torch.Generator(device=dev); gen.manual_seed(4)
ang = (torch.rand((B, T), generator=gen, device=dev, dtype=torch.float64) * (2 * np.pi))[lo:hi].contiguous()
tgt = np.full(16, 0.25)
for _ in range(3):
    g = psr_gradient_batched(ps, ang, tgt)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    g = psr_gradient_batched(ps, ang, tgt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    s = float(t.item()) * 1e-3
    print(json.dumps({"config": "C4-distributed", "gpus": world, "base_vectors": B, "ms": s * 1e3, "gradients_per_s": B / s,
                      "pattern_evals_per_s": B * 2 * T / s, "finite": bool(torch.isfinite(g).all().item())}), flush=True)
if world > 1:
    dist.destroy_process_group()
