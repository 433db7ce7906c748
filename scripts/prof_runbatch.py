import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mentpy_b200 as mb
B, T = 65536, 10
h = torch.empty((B, T), dtype=torch.float64).pin_memory(); h.uniform_(0, 6.28)
gs = mb.templates.grid_cluster(2, 6)
ps = mb.PatternSimulator(gs, backend="cuda-sv")
for _ in range(5): ps.run_batch(h, copy=False)
t0 = time.perf_counter()
for _ in range(50): ps.run_batch(h, copy=False)
print("us/step", (time.perf_counter() - t0) / 50 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): ps.run_batch(h, copy=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
