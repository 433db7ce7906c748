"""Kernel-level timing (GPU box) of the shared-memory SV kernel (6 <= w <= 12) and, for
comparison, the register kernel at w = 5: grid_cluster(w-1, 4), default window = rows + 1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from perf_sv import time_kernel

for w in (5, 6, 7, 8, 10, 12):
    for B in (4096, 65536):
        us = time_kernel(("grid_cluster", [w - 1, 4]), B, reps=40)
        M = (w - 1) * 3
        amps = B * M * 2**w
        print(f"w={w:2d} B={B:6d}: {us:10.1f} us/launch  {B/us:9.3f} M evals/s  {amps/us/1e3:8.2f} G amplitude-updates/s", flush=True)
