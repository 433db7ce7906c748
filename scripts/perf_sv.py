"""Kernel-level timing probe (GPU box): batched SV kernel time vs batch size, CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mentpy_b200 as mb
from mentpy_b200 import _lib

def time_kernel(spec, B, reps=200, form=_lib.OUT_SV):
    name, args = spec
    gs = getattr(mb.templates, name)(*args)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    sim = ps.simulator
    T, k = len(gs.trainable_nodes), len(gs.output_nodes)
    dev = torch.device("cuda")
    nbuf = max(2, int(200e6 // (B * (8 * T + 16 * 2**k))))
    nbuf = min(nbuf, 64)
    ang = torch.rand((nbuf, B, T), device=dev, dtype=torch.float64) * 6.28
    out = torch.empty((nbuf, B, 2**k if form == _lib.OUT_SV else 4**k), dtype=torch.complex128, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    lib = _lib.load(); plan = sim._full_plan()
    stream = torch.cuda.current_stream()
    ap = [ang[j].data_ptr() for j in range(nbuf)]; op = [out[j].data_ptr() for j in range(nbuf)]
    def go(j, s): 
        rc = lib.mbqc_run_batch_sv(plan.handle, ap[j], T, None, 0, B, op[j], form, st.data_ptr(), s)
        assert rc == 0
    for j in range(nbuf): go(j, stream.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    S = int(os.environ.get("PERF_STREAMS", "1"))
    side = [torch.cuda.Stream() for _ in range(S - 1)]
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        for st_ in side: st_.wait_stream(cur)
        for j in range(nbuf):
            s_ = cur if j % S == 0 else side[j % S - 1]
            go(j, s_.cuda_stream)
        for st_ in side: cur.wait_stream(st_)
    g.replay(); torch.cuda.synchronize()
    n = max(1, reps // nbuf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (n * nbuf)
    return us

if __name__ == "__main__":
    if len(sys.argv) > 1:  # e.g. perf_sv.py grid_cluster 2,6 1048576  (single config, for ncu)
        spec = (sys.argv[1], [int(x) for x in sys.argv[2].split(",")])
        B = int(sys.argv[3])
        print(spec, B, time_kernel(spec, B, reps=20))
        sys.exit(0)
    for spec in (("grid_cluster", [2, 6]), ("grid_cluster", [4, 5]), ("linear_cluster", [5])):
        for B in (1024, 8192, 32768, 65536, 262144, 1048576):
            us = time_kernel(spec, B)
            print(f"{spec} B={B:8d}  {us:9.2f} us/launch  {B/us/1e3:8.2f} G evals/s", flush=True)
