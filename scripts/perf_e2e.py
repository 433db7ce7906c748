"""Host end-to-end probe (GPU box): raw pinned PCIe bandwidth and the chunked pipeline."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mentpy_b200 as mb
from mentpy_b200 import _lib

B, T, K = 65536, 10, 4
dev = torch.device("cuda")
h = torch.empty((B, T), dtype=torch.float64).pin_memory(); h.uniform_(0, 6.28)
d = torch.empty((B, T), dtype=torch.float64, device=dev)
ho = torch.empty((B, K), dtype=torch.complex128).pin_memory()
do = torch.empty((B, K), dtype=torch.complex128, device=dev)
def bw(fn, nbytes, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    return nbytes / dt / 1e9, dt * 1e6
print("H2D 5.2MB: %.1f GB/s %.1f us" % bw(lambda: d.copy_(h, non_blocking=True), h.numel() * 8))
print("D2H 4.2MB: %.1f GB/s %.1f us" % bw(lambda: ho.copy_(do, non_blocking=True), ho.numel() * 16))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
print("H2D+D2H concurrent: %.1f GB/s %.1f us" % bw(both, h.numel() * 8 + ho.numel() * 16))

gs = mb.templates.grid_cluster(2, 6)
ps = mb.PatternSimulator(gs, backend="cuda-sv"); sim = ps.simulator
lib = _lib.load(); plan = sim._full_plan()
need = lib.mbqc_host_workspace_bytes(plan.handle, B, 0)
work = torch.empty(need, dtype=torch.uint8, device=dev)
flag = C.c_int32(0)
for chunks in (1, 2, 4, 6, 8, 12, 16, 32):
    def run():
        rc = lib.mbqc_run_batch_sv_host(plan.handle, h.data_ptr(), T, None, 0, B, ho.data_ptr(), 0, work.data_ptr(), need, C.byref(flag), chunks)
        assert rc == 0
    for _ in range(5): run()
    t0 = time.perf_counter()
    for _ in range(50): run()
    dt = (time.perf_counter() - t0) / 50
    print(f"C pipeline chunks={chunks:3d}: {dt*1e6:8.1f} us/step  {B/dt/1e6:8.1f} M evals/s")
t0 = time.perf_counter()
for _ in range(50): ps.run_batch(h, copy=False)
dt = (time.perf_counter() - t0) / 50
print(f"python run_batch(copy=False): {dt*1e6:8.1f} us/step  {B/dt/1e6:8.1f} M evals/s")

# same C call as run_batch makes (shared input state instead of the built-in |+> seed)
inp, mode = sim._stage_inputs(None, B, dev)
torch.cuda.synchronize()
for label, iptr, imode in (("INPUT_PLUS", None, 0), ("INPUT_SHARED", inp.data_ptr(), mode)):
    def run2():
        rc = lib.mbqc_run_batch_sv_host(plan.handle, h.data_ptr(), T, iptr, imode, B, ho.data_ptr(), 0, work.data_ptr(), need, C.byref(flag), 0)
        assert rc == 0
    for _ in range(5): run2()
    t0 = time.perf_counter()
    for _ in range(50): run2()
    dt = (time.perf_counter() - t0) / 50
    print(f"C pipeline default chunks, {label}: {dt*1e6:8.1f} us/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): ps.run_batch(h, copy=False)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(6)
