"""All BASELINE.json configs on one GPU (GPU box): device-resident throughput with CUDA events.
C1 linear_cluster(5) SV, C2 grid 2x6 SV, C3 grid 3x8 DM (+depolarizing), C4 grid 4x5 psr gradient,
C5 streaming linear_cluster(w+16, window w).  One JSON line per config -> profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import mentpy_b200 as mb
from mentpy_b200 import _lib
from mentpy_b200.gradients import psr_gradient_batched

PEAK = 6459.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps

def emit(**kw):
    print(json.dumps(kw), flush=True)

dev = torch.device("cuda")
# C1 / C2: batched SV
for tag, spec, B in (("C1", ("linear_cluster", [5]), 1 << 20), ("C2", ("grid_cluster", [2, 6]), 65536), ("C2-large-batch", ("grid_cluster", [2, 6]), 1 << 22)):
    gs = getattr(mb.templates, spec[0])(*spec[1]); T, k = len(gs.trainable_nodes), len(gs.output_nodes)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    ang = torch.rand((B, T), device=dev, dtype=torch.float64) * 6.283
    t = timeit(lambda: ps.run_batch(ang), 50)
    by = B * (8 * T + 16 * 2**k)
    emit(config=tag, pattern=f"{spec[0]}{spec[1]}", batch=B, evals_per_s=B / t, ms=t * 1e3, algorithmic_GBps=by / t / 1e9, hbm_frac=by / t / 1e9 / PEAK, note="single stream, torch out alloc inside")
# C2 in complex64 mode
gs = mb.templates.grid_cluster(2, 6); T, k = 10, 2
for B in (65536, 1 << 22):
    ps32 = mb.PatternSimulator(gs, backend="cuda-sv", dtype="complex64")
    ang = torch.rand((B, T), device=dev, dtype=torch.float64) * 6.283
    t = timeit(lambda: ps32.run_batch(ang), 50)
    by = B * (8 * T + 8 * 2**k)
    emit(config="C2-complex64", pattern="grid_cluster[2, 6]", batch=B, evals_per_s=B / t, ms=t * 1e3, algorithmic_GBps=by / t / 1e9, hbm_frac=by / t / 1e9 / PEAK)
# C3: DM with noise
gs = mb.templates.grid_cluster(3, 8); T = len(gs.trainable_nodes)
for p in (0.0, 0.01, 0.1):
    ps = mb.PatternSimulator(gs, backend="cuda-dm", **({} if p == 0 else {"circuit_noise": "depolarizing", "p": p}))
    ang = torch.rand((4096, T), device=dev, dtype=torch.float64) * 6.283
    t = timeit(lambda: ps.run_batch(ang), 20)
    by = 4096 * (8 * T + 16 * 64)
    emit(config="C3", pattern="grid_cluster(3,8) DM", depolarizing_p=p, batch=4096, evals_per_s=4096 / t, ms=t * 1e3, algorithmic_GBps=by / t / 1e9, hbm_frac=by / t / 1e9 / PEAK)
# C3 at kernel level (C-ABI launches captured in a CUDA graph, no Python per-call overhead)
from perf_dm import time_kernel as dm_time_kernel  # noqa: E402
for p, B in ((0.0, 4096), (0.01, 4096), (0.0, 65536), (0.01, 65536)):
    us = dm_time_kernel([3, 8], B, p)
    by = B * (8 * T + 16 * 64)
    emit(config="C3-kernel", pattern="grid_cluster(3,8) DM", depolarizing_p=p, batch=B, evals_per_s=B / us * 1e6, us_per_launch=us,
         algorithmic_GBps=by / us / 1e3, hbm_frac=by / us / 1e3 / PEAK)
# sampled runs (force0=False): Philox outcomes + flow corrections
gs = mb.templates.grid_cluster(2, 6); T = len(gs.trainable_nodes)
ps = mb.PatternSimulator(gs, backend="cuda-sv", force0=False, seed=1)
ang = torch.rand((1 << 20, T), device=dev, dtype=torch.float64) * 6.283
t = timeit(lambda: ps.sample_batch(ang), 20)
emit(config="C2-sampled", pattern="grid_cluster(2,6) SV force0=False", batch=1 << 20, shots_per_s=(1 << 20) / t, ms=t * 1e3)
gs = mb.templates.grid_cluster(3, 8); T = len(gs.trainable_nodes)
ps = mb.PatternSimulator(gs, backend="cuda-dm", force0=False, seed=1, circuit_noise="depolarizing", p=0.01)
ang = torch.rand((1 << 16, T), device=dev, dtype=torch.float64) * 6.283
t = timeit(lambda: ps.sample_batch(ang), 10)
emit(config="C3-sampled", pattern="grid_cluster(3,8) DM force0=False depolarizing 0.01", batch=1 << 16, shots_per_s=(1 << 16) / t, ms=t * 1e3)
# C4: gradient
gs = mb.templates.grid_cluster(4, 5); T = len(gs.trainable_nodes)
ps = mb.PatternSimulator(gs, backend="cuda-sv")
tgt = np.full(16, 0.25)
for B in (1 << 16, 1 << 20):
    ang = torch.rand((B, T), device=dev, dtype=torch.float64) * 6.283
    t = timeit(lambda: psr_gradient_batched(ps, ang, tgt), 5)
    emit(config="C4", pattern="grid_cluster(4,5) psr gradient", base_vectors=B, gradients_per_s=B / t, pattern_evals_per_s=B * 2 * T / t, ms=t * 1e3, algorithmic_GBps=B * 256 / t / 1e9, hbm_frac=B * 256 / t / 1e9 / PEAK)
# data-set averaged gradient (one optimiser step of the QML tutorial): P vectors x S data items
from mentpy_b200.gradients import psr_gradient_dataset  # noqa: E402
P, S = 64, 4096
X = torch.rand((P, T), device=dev, dtype=torch.float64) * 6.283
ins = torch.randn((S, 16), device=dev, dtype=torch.complex128); ins = ins / ins.norm(dim=1, keepdim=True)
tgs = torch.randn((S, 16), device=dev, dtype=torch.complex128); tgs = tgs / tgs.norm(dim=1, keepdim=True)
t = timeit(lambda: psr_gradient_dataset(ps, X, tgs, ins), 5)
emit(config="C4-dataset", pattern="grid_cluster(4,5) data-set psr gradient", vectors=P, data_items=S,
     gradients_per_s=P / t, pattern_evals_per_s=P * S * 2 * T / t, ms=t * 1e3)
# C5: streaming
for w in (28, 30, 32):
    try:
        gs = mb.templates.linear_cluster(w + 16)
        ps = mb.PatternSimulator(gs, backend="cuda-sv-stream", window_size=w, fuse=5)
        ang = np.random.default_rng(4).uniform(0, 2 * np.pi, w + 15)
        ps.run(ang)
        t = 1e9
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter(); ps.run(ang); torch.cuda.synchronize(); t = min(t, time.perf_counter() - t0)
        s = ps.simulator.last_schedule
        emit(config="C5", pattern=f"linear_cluster({w+16}) window {w}", state_GiB=16 * 2**w / 2**30, fuse=5, s_per_pattern=t, passes=len(s.passes),
             algorithmic_GBps=s.algorithmic_bytes / t / 1e9, hbm_frac_algorithmic=s.algorithmic_bytes / t / 1e9 / PEAK, streamed_GBps=s.streamed_bytes / t / 1e9)
        del ps
    except Exception as e:
        emit(config="C5", window=w, error=repr(e)[:200])
# device-resident training loop (mbqc_train_dataset): the tutorial workload and a wide one
from mentpy_b200.optimizers import adam_optimize_batched  # noqa: E402
gt = mb.templates.muta(2, 1, one_column=True); gt[3] = mb.Ment("X"); gt[8] = mb.Ment("X")
pt = mb.PatternSimulator(gt, backend="cuda-sv"); Tt = len(gt.trainable_nodes)
for P, S, iters in ((1, 7, 100), (1024, 64, 100)):
    X = torch.rand((P, Tt), device=dev, dtype=torch.float64) * 6.283
    ins = torch.randn((S, 4), device=dev, dtype=torch.complex128); ins = ins / ins.norm(dim=1, keepdim=True)
    tgs = torch.randn((S, 4), device=dev, dtype=torch.complex128); tgs = tgs / tgs.norm(dim=1, keepdim=True)
    for fused in (True, False):
        def go():
            adam_optimize_batched(pt, X, tgs, num_iters=iters, step_size=0.08, input_states=ins, dataset=True, fused=fused)
            torch.cuda.synchronize()
        go()
        t0 = time.perf_counter(); go(); t = time.perf_counter() - t0
        emit(config="train-loop", pattern="muta(2,1,one_column) X on 3,8 (intro-to-mbqml.rst)", vectors=P, data_items=S, iterations=iters,
             fused_c_loop=fused, s_total=t, us_per_iteration=t / iters * 1e6, pattern_evals_per_s=P * S * 2 * Tt * iters / t)
