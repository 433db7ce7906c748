"""Host end-to-end probe (GPU box): submit/wait pipeline at the C ABI, sweep of calls in flight and
chunks per call; then the Python run_batch_async path."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mentpy_b200 as mb
from mentpy_b200 import _lib

B, T, K = 65536, 10, 4
dev = torch.device("cuda")
NB = 4
h = [torch.empty((B, T), dtype=torch.float64).pin_memory() for _ in range(NB)]
for x in h: x.uniform_(0, 6.28)
ho = [torch.empty((B, K), dtype=torch.complex128).pin_memory() for _ in range(NB)]
gs = mb.templates.grid_cluster(2, 6)
ps = mb.PatternSimulator(gs, backend="cuda-sv"); sim = ps.simulator
lib = _lib.load(); plan = sim._full_plan()
need = lib.mbqc_host_workspace_bytes(plan.handle, B, 0)
work = [torch.empty(need, dtype=torch.uint8, device=dev) for _ in range(NB)]
flag = C.c_int32(0); tk = C.c_int32(0)

def run(steps, depth, chunks):
    pend = []
    for i in range(steps):
        j = i % NB
        rc = lib.mbqc_run_batch_sv_host_submit(plan.handle, h[j].data_ptr(), T, None, 0, B, ho[j].data_ptr(), 0,
                                               work[j].data_ptr(), need, chunks, C.byref(tk))
        assert rc == 0, lib.mbqc_last_error()
        pend.append(tk.value)
        if len(pend) == depth:
            assert lib.mbqc_host_wait(pend.pop(0), C.byref(flag)) == 0
    for t in pend:
        assert lib.mbqc_host_wait(t, C.byref(flag)) == 0

for depth in (1, 2, 3):
    for chunks in (1, 2, 4, 8):
        run(6, depth, chunks)
        t0 = time.perf_counter(); run(60, depth, chunks); dt = (time.perf_counter() - t0) / 60
        print(f"C submit/wait depth={depth} chunks={chunks}: {dt*1e6:7.1f} us/step {B/dt/1e6:7.1f} M evals/s", flush=True)

def pipelined(n):
    pend = []
    for i in range(n):
        pend.append(ps.run_batch_async(h[i % NB]))
        if len(pend) == 2: pend.pop(0).result()
    for p in pend: p.result()
pipelined(6)
t0 = time.perf_counter(); pipelined(60); dt = (time.perf_counter() - t0) / 60
print(f"python run_batch_async depth 2: {dt*1e6:7.1f} us/step {B/dt/1e6:7.1f} M evals/s")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); pipelined(60); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
