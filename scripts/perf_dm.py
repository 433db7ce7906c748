"""Kernel-level timing probe (GPU box): batched DM kernel time vs batch size through the C-ABI,
CUDA graph of several launches on distinct buffers, CUDA events.  `perf_dm.py 3,8 4096 [p]` runs a
single config (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mentpy_b200 as mb
from mentpy_b200 import _lib


def time_kernel(shape, B, p=0.0, reps=200, window=None):
    gs = mb.templates.grid_cluster(*shape)
    kw = {} if p == 0 else {"circuit_noise": "depolarizing", "p": p}
    if window:
        kw["window_size"] = window
    ps = mb.PatternSimulator(gs, backend="cuda-dm", **kw)
    sim = ps.simulator
    T, k = len(gs.trainable_nodes), len(gs.output_nodes)
    dev = torch.device("cuda")
    nbuf = min(32, max(2, int(300e6 // (B * (8 * T + 16 * 4**k)))))
    ang = torch.rand((nbuf, B, T), device=dev, dtype=torch.float64) * 6.28
    out = torch.empty((nbuf, B, 4**k), dtype=torch.complex128, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    lib = _lib.load(); plan = sim._full_plan()
    def go(j, s):
        rc = lib.mbqc_run_batch_dm(plan.handle, ang[j].data_ptr(), T, None, 0, B, out[j].data_ptr(), None, st.data_ptr(), s)
        assert rc == 0
    for j in range(nbuf): go(j, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for j in range(nbuf): go(j, torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    n = max(1, reps // nbuf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * nbuf)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] != "quick":
        shape = [int(x) for x in sys.argv[1].split(",")]
        B = int(sys.argv[2]); p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
        print(shape, B, p, time_kernel(shape, B, p, reps=20))
        sys.exit(0)
    import json
    lib = _lib.load()
    if len(sys.argv) > 1 and sys.argv[1] == "quick":  # specialised kernel only (MBQC_DM_JIT_LB picks the layout)
        lib.mbqc_jit_set_mode(2)
        for shape in ([3, 8], [4, 5]):
            for p in (0.0, 0.01):
                for B in (1024, 4096, 16384, 65536, 262144):
                    us = time_kernel(shape, B, p)
                    print(json.dumps({"pattern": f"grid_cluster{tuple(shape)} DM", "p": p, "batch": B, "lb": os.environ.get("MBQC_DM_JIT_LB", "default"),
                                      "mbqc_jit_dm_us": round(us, 2), "Mevals_s": round(B / us, 2)}), flush=True)
        sys.exit(0)
    for shape, window in (([3, 8], None), ([2, 6], None), ([4, 5], None), ([5, 4], None)):
        for p in (0.0, 0.01):
            for B in (1024, 4096, 16384, 65536, 262144):
                row = {"pattern": f"grid_cluster{tuple(shape)} DM", "p": p, "batch": B}
                for mode, name in ((0, "dm_reg_kernel"), (2, "mbqc_jit_dm")):  # general vs specialised kernel
                    lib.mbqc_jit_set_mode(mode)
                    us = time_kernel(shape, B, p, window=window)
                    row[name + "_us"] = round(us, 2)
                    row[name + "_Mevals_s"] = round(B / us, 2)
                row["speedup"] = round(row["dm_reg_kernel_us"] / row["mbqc_jit_dm_us"], 2)
                print(json.dumps(row), flush=True)
