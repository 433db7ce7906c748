"""Forward register kernel vs window (GPU box): throughput regime, per sample-step rates."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from perf_sv import time_kernel
for spec, w, M in ((("linear_cluster", [17]), 2, 16), (("grid_cluster", [2, 9]), 3, 16), (("grid_cluster", [3, 6]), 4, 15), (("grid_cluster", [4, 5]), 5, 16)):
    for B in (65536, 1 << 20):
        us = time_kernel(spec, B, reps=40)
        print(f"w={w} {spec} B={B:8d}: {us:9.1f} us  {B/us/1e3:7.2f} G evals/s  {B*M/us/1e3:7.1f} G sample-steps/s  {B*M*2**w/us/1e6:6.2f} T amp-updates/s", flush=True)
