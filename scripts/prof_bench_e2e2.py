import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mentpy_b200 as mb
B, T = 65536, 10
dev = torch.device("cuda")
gs = mb.templates.grid_cluster(2, 6)
ps = mb.PatternSimulator(gs, backend="cuda-sv")
def measure(label, bufs, n=50):
    for i in range(3): ps.run_batch(bufs[i % len(bufs)], copy=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): ps.run_batch(bufs[i % len(bufs)], copy=False)
    torch.cuda.synchronize(); print(f"{label}: {(time.perf_counter()-t0)/n*1e6:8.1f} us/step", flush=True)
h4 = [torch.empty((B, T), dtype=torch.float64).pin_memory().uniform_(0, 6.28) for _ in range(4)]
measure("early buffers, uniform(0,6.28)", h4)
big = torch.rand((4, B, T), device=dev, dtype=torch.float64) * 6.28
for j, h in enumerate(h4): h.copy_(big[j].cpu())
measure("same buffers refilled from device rand*6.28 via .cpu()", h4)
for h in h4: h.uniform_(0, 6.28)
measure("same buffers refilled with uniform_", h4)
for h in h4: h.copy_(torch.rand((B, T), dtype=torch.float64) * 6.28)
measure("same buffers refilled from a CPU rand tensor", h4)
for h in h4: h.uniform_(0, 1.0)
measure("same buffers uniform(0,1)", h4)
late = [torch.empty((B, T), dtype=torch.float64).pin_memory().uniform_(0, 6.28) for _ in range(4)]
measure("late buffers, uniform(0,6.28)", late)
