#!/usr/bin/env python
"""Benchmark of the MBQC pattern-evaluation hot path (BASELINE.json metric: pattern evals/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (N = 1 and per rank for N > 1, weak scaling): BASELINE.json configs[1] --
grid_cluster(2,6) state-vector pattern, 65,536 random angle sets per step, outputs [B,4]
complex128.  A step = one pass of the hot path over one batch = ONE kernel launch.  Batches rotate
through a pool whose angles+outputs exceed the 126 MB L2, so no step finds its inputs cached.

Prints ONE JSON line (rank 0):
  value / ms_per_step   device-resident throughput: the K-step block is timed `repeats` times
                        (CUDA events on the launching stream, barrier + synchronize on both sides of
                        every block, max over ranks per block) and the MEDIAN block is reported;
                        `first_block_ms` is the single block the base contract asks for
  e2e                   same metric through PatternSimulator.run_batch_async with HOST buffers (pinned
                        H2D + results into host memory inside the timed region); `host_floor_ms` is a
                        raw concurrent H2D + D2H copy of the same bytes on the same ranks, no kernels
  roofline              algorithmic bytes per launch / launch time vs the measured HBM peak, plus the
                        FP64 roof measured live with an FMA microbenchmark (roofline.fp64) and which
                        one binds (roofline.binds, backed by the ncu counters under profiles/)
  configs               the other BASELINE configs under the same clock: C1, C3 (p = 0 / 0.01), C4,
                        C5 at N = 1; at N > 1 C4 strong scaling with the all_gather inside the timed
                        region and C5 sharded (strong w = 32, weak w = 32 + log2 N) over NVLink peer
                        memory -- each with an in-run parity figure
  cpu_baseline          the dense numpy port of the reference (oracle/dense_port.py) on host cores
`--impl reference` times only that CPU port (all host cores), same metric / config.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ROWS, COLS = 2, 6
BATCH = 65536
N_ANGLES, N_OUT = 10, 2
SEED = 1
ALGO_BYTES_PER_EVAL = 8 * N_ANGLES + 16 * 2**N_OUT  # 8*T angles in + 16*2^k amplitudes out = 144 B
L2_BYTES = 126 * 1024 * 1024
PER_BATCH_BYTES = BATCH * ALGO_BYTES_PER_EVAL
POOL = max(4, int(np.ceil(1.5 * L2_BYTES / PER_BATCH_BYTES)))  # 21 batches = 189 MiB
GRAPH_BRANCHES = 8
REF_EVALS_PER_STEP = 8192  # reference arm: bounded sample of the 65,536-set batch per step


def bench_config(world):
    """`config` of the JSON line: the workload definition, identical in both arms."""
    return {
        "workload": "grid_cluster(2,6) statevector, 65,536 random angle sets per step per GPU (BASELINE configs[1])",
        "pattern": "grid_cluster(2,6)", "batch_per_gpu": BATCH, "window": 3, "measurements": 10,
        "output": "sv [B,4] complex128", "angles": "uniform [0, 2 pi), float64",
        "l2": f"no step re-reads a cached batch: inputs rotate through a pool of {POOL} batches "
              f"({POOL * PER_BATCH_BYTES / 2**20:.0f} MiB > 126 MiB L2); the CPU arm draws fresh angle sets every step",
        "parallelism": f"batch-split x{world}, no collective on the data path",
    }


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return None


def measured_peak_gbs():
    d = load_json("MEASURED_PEAKS.json")
    if d and "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_facts():
    """Per-launch facts of the dominant kernel from the committed ncu capture (profiles/traffic.json)."""
    return load_json("profiles", "traffic.json") or {}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the dense port of the reference, one process per host core
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    pat_json, angles = args
    from oracle.dense_port import DensePatternSV
    from oracle.pattern_data import PatternData

    pat = PatternData.from_json(pat_json)
    sim = DensePatternSV(pat)
    acc = 0.0
    for a in angles:
        sim.reset()
        acc += float(np.abs(sim.run(a, output_form="sv")[0]))
    return acc


def _pattern_json():
    import mentpy_b200 as mb
    from oracle.pattern_data import PatternData

    return PatternData.from_circuit(mb.templates.grid_cluster(ROWS, COLS)).to_json()


def cpu_port_rate(evals_per_core, cores):
    """evals/s of the dense reference port using `cores` processes (each single-threaded)."""
    import multiprocessing as mp

    pat_json = _pattern_json()
    rng = np.random.default_rng(SEED)
    chunks = [rng.uniform(0, 2 * np.pi, (evals_per_core, N_ANGLES)) for _ in range(cores)]
    pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None
    try:
        if pool is not None:
            pool.map(_cpu_worker, [(pat_json, c[:2]) for c in chunks])  # warm the workers
        t0 = time.perf_counter()
        if pool is not None:
            pool.map(_cpu_worker, [(pat_json, c) for c in chunks])
        else:
            _cpu_worker((pat_json, chunks[0]))
        dt = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    return evals_per_core * cores / dt, dt


def run_reference_arm(args):
    """--impl reference: the reference algorithm's CPU port on all host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    os.environ["OMP_NUM_THREADS"] = "1"
    cores = os.cpu_count() or 1
    per_core = max(1, REF_EVALS_PER_STEP // cores)
    sample = per_core * cores
    pat_json = _pattern_json()
    rng = np.random.default_rng(SEED)
    pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None

    def step():
        chunks = [(pat_json, rng.uniform(0, 2 * np.pi, (per_core, N_ANGLES))) for _ in range(cores)]
        if pool is not None:
            pool.map(_cpu_worker, chunks)
        else:
            _cpu_worker(chunks[0])

    steps, warmup = args.steps, args.warmup
    # keep the whole run within a few minutes whatever K is: ~1 s per step on 16 cores
    budget_steps = 150
    timed = min(steps, budget_steps)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(timed):
        step()
    dt = time.perf_counter() - t0
    if pool is not None:
        pool.close()
        pool.join()
    value = sample * timed / dt
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": "pattern_evals_per_s", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * dt / timed, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(max(world, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} angle sets of the 65,536-set batch per step x {timed} timed steps"
                                   + ("" if timed == steps else f" (of the {steps} requested: bounded run)")
                                   + ", one single-threaded process per core (oracle/dense_port.py restates "
                                     "np_simulator_sv.py incl. its dense kron operators; the live reference measured 1.76x slower)"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample_once(self):
        if not self.ok:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample_once()
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def pin_to_gpu_numa_node(index):
    """Bind this rank (and the page-locked buffers it allocates afterwards) to the CPU cores NVML
    reports as local to its GPU: the host<->device legs then stay on the GPU's own PCIe root."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        pass
    return 0


# ------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """Rank / device / collective helpers shared by the legs of the GPU arm."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa_cores = pin_to_gpu_numa_node(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        """element-wise max over ranks of a list of floats"""
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def gather_per_rank(self, value):
        if self.world == 1:
            return [float(value)]
        t = self.torch.zeros(self.world, dtype=self.torch.float64, device=self.dev)
        t[self.rank] = value
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]


def timed_blocks(ctx, launch_block, repeats):
    """Time `launch_block()` (enqueues one K-step block on the current stream) `repeats` times;
    every block is bracketed by barrier + synchronize; returns per-block ms, max over ranks."""
    torch = ctx.torch
    stream = torch.cuda.current_stream(ctx.dev)
    ms = []
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        e0.record(stream)
        launch_block()
        e1.record(stream)
        ctx.barrier()
        ms.append(e0.elapsed_time(e1))
    return ctx.max_over_ranks(ms)


def graph_of(ctx, launches, branches=1):
    """Capture `launches` (callables taking a raw cudaStream_t) as independent nodes spread over
    `branches` parallel branches of one CUDA graph."""
    torch = ctx.torch
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(device=ctx.dev)
    side = [torch.cuda.Stream(device=ctx.dev) for _ in range(max(branches, 1) - 1)]
    with torch.cuda.graph(g, stream=cap):
        cur = torch.cuda.current_stream(ctx.dev)
        for sd in side:
            sd.wait_stream(cur)
        for j, fn in enumerate(launches):
            br = j % max(branches, 1)
            fn((cur if br == 0 else side[br - 1]).cuda_stream)
        for sd in side:
            cur.wait_stream(sd)
    return g


def pool_size(bytes_per_batch, cap=64):
    return int(min(cap, max(2, np.ceil(1.5 * L2_BYTES / bytes_per_batch))))


def golden_case(pattern, backend, form):
    d = load_json("tests", "golden", "sim_cases.json")
    if not d:
        return None
    for c in d["cases"]:
        if c["spec"][0] == pattern[0] and c["spec"][1] == pattern[1] and not c["spec"][2] and c["backend"] == backend \
                and c["output_form"] == form and not c["x_nodes"] and not c["fixed"]:
            return c
    return None


def cplx(d):
    return (np.asarray(d["re"], dtype=float) + 1j * np.asarray(d["im"], dtype=float)).reshape(d["shape"])


def linear_cluster_closed_form(angles):
    """Output of linear_cluster(L) on |+>: J(-th_{L-2}) ... J(-th_0)|+>, J(a) = [[1, e^{ia}], [1, -e^{ia}]]/sqrt2,
    for any window size (SURVEY.md 8c 'analytic oracle').  A closed form, not the oracle package."""
    v = np.array([1.0, 1.0], dtype=complex) / np.sqrt(2.0)
    for th in np.asarray(angles, dtype=float):
        e = np.exp(-1j * th)
        v = np.array([v[0] + e * v[1], v[0] - e * v[1]]) / np.sqrt(2.0)
    return v / np.linalg.norm(v)


def infidelity(a, b):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    return float(abs(1.0 - abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real)))


# ------------------------------------------------------------------------------------------------
# second roofline: FP64 FMA peak, measured live
# ------------------------------------------------------------------------------------------------
def fp64_peak_tflops(ctx):
    import ctypes as C

    from mentpy_b200 import _lib

    torch = ctx.torch
    lib = _lib.load()
    out = torch.zeros(8, dtype=torch.float64, device=ctx.dev)
    flops = C.c_int64()
    stream = torch.cuda.current_stream(ctx.dev)
    best = None
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(lib.mbqc_probe_fp64_fma(8192, 148 * 8, 256, out.data_ptr(), C.byref(flops), stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize(ctx.dev)
        if i >= 1:
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
    return flops.value / (best * 1e-3) / 1e12


# ------------------------------------------------------------------------------------------------
# host floor: the e2e step's bytes as raw concurrent copies, no kernels
# ------------------------------------------------------------------------------------------------
def host_copy_floor(ctx, h2d_bytes, d2h_bytes, steps=40):
    """ms per step of the step's copies alone: H2D only, D2H only, and both directions at once
    (two pinned DMAs per step on two streams), each the best of 3 runs, max over ranks."""
    torch = ctx.torch
    hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device=ctx.dev)
    dout = torch.empty(d2h_bytes, dtype=torch.uint8, device=ctx.dev)
    s_in, s_out = torch.cuda.Stream(device=ctx.dev), torch.cuda.Stream(device=ctx.dev)

    def go(n, up, down):
        for _ in range(n):
            if up:
                with torch.cuda.stream(s_in):
                    din.copy_(hin, non_blocking=True)
            if down:
                with torch.cuda.stream(s_out):
                    hout.copy_(dout, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    res = {}
    for name, up, down in (("h2d_only_ms", True, False), ("d2h_only_ms", False, True), ("both_directions_ms", True, True)):
        go(3, up, down)
        runs = []
        for _ in range(3):
            ctx.barrier()
            t0 = time.perf_counter()
            go(steps, up, down)
            runs.append(time.perf_counter() - t0)
        res[name] = 1e3 * min(ctx.max_over_ranks(runs)) / steps
    return res


# ------------------------------------------------------------------------------------------------
# the other BASELINE configs (device-resident, CUDA events, max over ranks)
# ------------------------------------------------------------------------------------------------
def time_rotating(ctx, make_call, n_buf, reps):
    """make_call(j) -> callable launching batch j on the current stream (returns its output).
    One CUDA graph holds a pass over the n_buf batches; `reps` replays are timed as one block."""
    torch = ctx.torch
    outs = [make_call(j)() for j in range(n_buf)]  # warm-up, also sizes the allocator
    torch.cuda.synchronize(ctx.dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = [make_call(j)() for j in range(n_buf)]
    g.replay()
    torch.cuda.synchronize(ctx.dev)
    ms = timed_blocks(ctx, lambda: [g.replay() for _ in range(reps)], 5)
    return float(np.median(ms)) / (reps * n_buf), outs


# SURVEY 8d flop model of the DENSE algorithm per unit of work: SV M (8 2^w + c_sincos), DM M 10 4^w
ALGO_FLOPS = {"C1": (4 * (8 * 4 + 70), "evaluation"), "C3": (55e3, "evaluation"), "C4": (154e3, "gradient (32 evaluations)")}


def annotate_fp64(configs, fp64_peak, world):
    """Second roof of the register-resident configs: algorithmic flops x rate against the
    measured DFMA peak.  The kernels execute FEWER flops than the dense model (compressed state, linear
    split of the shift rule), so the algorithmic fraction can exceed what the pipe counters show."""
    for name, (flops, unit) in ALGO_FLOPS.items():
        legs = [configs.get(name)] if name != "C3" else [configs.get("C3", {}).get("p0"), configs.get("C3", {}).get("depolarizing_p0.01"), configs.get("C3")]
        for c in legs:
            if not c or "value" not in c or "roofline" not in c:
                continue
            tf = c["value"] / world * flops / 1e12
            c["roofline"]["fp64_algorithmic"] = {"flops_per_unit": flops, "unit": unit, "achieved_tflops": tf,
                                                 "peak_measured_tflops": fp64_peak, "frac": tf / fp64_peak,
                                                 "note": "per GPU; SURVEY 8d flop model of the dense algorithm"}


def config_c1(ctx, peak):
    import mentpy_b200 as mb

    torch = ctx.torch
    B, T, k = 1 << 20, 4, 1
    gs = mb.templates.linear_cluster(5)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    per = B * (8 * T + 16 * 2**k)
    n_buf = pool_size(per)
    gen = torch.Generator(device=ctx.dev)
    gen.manual_seed(1000 * ctx.rank)
    ang = torch.rand((n_buf, B, T), generator=gen, device=ctx.dev, dtype=torch.float64) * (2 * np.pi)
    gold = golden_case(("linear_cluster", [5]), "numpy-sv", "sv")
    if gold:
        ang[0, 0] = torch.tensor(gold["angles"], dtype=torch.float64)
    ms, outs = time_rotating(ctx, lambda j: (lambda: ps.run_batch(ang[j])), n_buf, 8)
    res = {"pattern": "linear_cluster(5) SV (BASELINE configs[0])", "batch": B, "window": 2, "measurements": 4,
           "value": B / (ms * 1e-3), "unit": "evals/s", "ms": ms,
           "roofline": {"bound": "hbm", "achieved": per / (ms * 1e-3) / 1e9, "frac": per / (ms * 1e-3) / 1e9 / peak,
                        "algorithmic_bytes_per_launch": per},
           "l2": f"{n_buf} batches of {per / 2**20:.0f} MiB rotate (> L2)"}
    if gold:
        got = outs[0][0].cpu().numpy()
        res["parity"] = {"against": "tests/golden/sim_cases.json (recorded from the reference), row 0",
                         "infidelity": infidelity(got, cplx(gold["output"])),
                         "max_abs_diff": float(np.abs(got - cplx(gold["output"])).max())}
    return res


def config_c3(ctx, peak):
    import mentpy_b200 as mb

    torch = ctx.torch
    B, k = 4096, 3
    gs = mb.templates.grid_cluster(3, 8)
    T = len(gs.trainable_nodes)
    per = B * (8 * T + 16 * 4**k)
    n_buf = pool_size(per)
    gen = torch.Generator(device=ctx.dev)
    gen.manual_seed(2 + 1000 * ctx.rank)
    ang = torch.rand((n_buf, B, T), generator=gen, device=ctx.dev, dtype=torch.float64) * (2 * np.pi)
    gold = golden_case(("grid_cluster", [3, 8]), "numpy-dm", "dm")
    if gold:
        ang[0, 0] = torch.tensor(gold["angles"], dtype=torch.float64)
    res = {"pattern": "grid_cluster(3,8) density matrix (BASELINE configs[2])", "batch": B, "window": 4,
           "measurements": 21, "unit": "evals/s", "l2": f"{n_buf} batches of {per / 2**20:.1f} MiB rotate (> L2)"}
    for p in (0.0, 0.01):
        kw = {} if p == 0 else {"circuit_noise": "depolarizing", "p": p}
        ps = mb.PatternSimulator(gs, backend="cuda-dm", **kw)
        ms, outs = time_rotating(ctx, lambda j: (lambda: ps.run_batch(ang[j])), n_buf, 4)
        tag = "p0" if p == 0 else "depolarizing_p0.01"
        res[tag] = {"value": B / (ms * 1e-3), "ms": ms,
                    "roofline": {"bound": "hbm", "achieved": per / (ms * 1e-3) / 1e9,
                                 "frac": per / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": per}}
        rho = outs[0][0].cpu().numpy()
        if p == 0 and gold:
            res[tag]["parity"] = {"against": "tests/golden/sim_cases.json (recorded from the reference), row 0",
                                  "max_abs_diff": float(np.abs(rho - cplx(gold["output"])).max())}
        else:
            res[tag]["parity"] = {"against": "invariants (noise parity is unpinned: the reference has no numpy noise path)",
                                  "trace_minus_1": float(abs(np.trace(rho).real - 1.0)),
                                  "hermiticity": float(np.abs(rho - rho.conj().T).max())}
    res["value"] = res["p0"]["value"]
    res["ms"] = res["p0"]["ms"]
    res["roofline"] = res["p0"]["roofline"]
    return res


def gather_path(ps):
    """How the replicated result of the last psr_gradient_distributed call travels."""
    cache = getattr(getattr(ps, "simulator", ps), "_replicated_results", {}) or {}
    mc = any(getattr(r, "mc_ptr", 0) for r in cache.values())
    return ("multimem stores to ONE NVSwitch multicast address (torch symmetric memory): the switch replicates them, 1/N of the "
            "per-peer NVLink egress" if mc else "NVLink peer stores / bulk copies into CUDA-IPC mapped copies")


def config_c4(ctx, peak):
    """2^20 angle vectors, parameter-shift gradient of grid_cluster(4,5); at N > 1 the vectors are
    split across the ranks (strong scaling) and the gather is INSIDE the timed region: the gradient
    kernel stores its rows into every GPU's copy of the result over NVLink peer memory, one flag
    barrier closes the step (dist.psr_gradient_distributed); the NCCL all_gather form is timed beside it."""
    import mentpy_b200 as mb
    from mentpy_b200.dist import psr_gradient_distributed, slice_bounds
    from mentpy_b200.gradients import psr_gradient_batched

    torch = ctx.torch
    B = 1 << 20
    gs = mb.templates.grid_cluster(4, 5)
    T = len(gs.trainable_nodes)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    gen = torch.Generator(device=ctx.dev)
    gen.manual_seed(4)  # same stream on every rank: the ranks hold slices of ONE batch
    full = torch.rand((B, T), generator=gen, device=ctx.dev, dtype=torch.float64) * (2 * np.pi)
    gold = load_json("tests", "golden", "gradients.json")
    gold = gold["c4"] if gold and "c4" in gold else None
    tgt_np = cplx(gold["target"]).reshape(-1) if gold else np.full(16, 0.25, dtype=complex)
    if gold:
        full[0] = torch.tensor(gold["x"], dtype=torch.float64)
    tgt = torch.as_tensor(tgt_np).to(ctx.dev)
    lo, hi = slice_bounds(B, ctx.rank, ctx.world)
    part = full[lo:hi]

    def step():
        return psr_gradient_distributed(ps, full, tgt) if ctx.world > 1 else psr_gradient_batched(ps, part, tgt)

    def step_nccl():
        return psr_gradient_distributed(ps, full, tgt, fused=False)

    for _ in range(2):
        grad = step()
    reps = 5
    ms = float(np.median(timed_blocks(ctx, lambda: [step() for _ in range(reps)], 5))) / reps
    ms_nccl = None
    if ctx.world > 1:
        for _ in range(2):
            gn = step_nccl()
        ms_nccl = float(np.median(timed_blocks(ctx, lambda: [step_nccl() for _ in range(reps)], 5))) / reps
        grad = step()
        agree = float((gn - grad).abs().max().item())
        del gn
    grad = step()
    # compute-only time of this rank's slice, for the record
    ms_local = float(np.median(timed_blocks(ctx, lambda: [psr_gradient_batched(ps, part, tgt) for _ in range(reps)], 3))) / reps
    per = B * 256
    res = {"pattern": "grid_cluster(4,5) parameter-shift gradient, 2^20 angle vectors (BASELINE configs[3])",
           "base_vectors": B, "window": 5, "measurements": 16, "evals_per_gradient": 2 * T,
           "value": B / (ms * 1e-3), "unit": "gradients/s", "pattern_evals_per_s": B * 2 * T / (ms * 1e-3), "ms": ms,
           "ms_compute_only": ms_local, "scaling": "strong" if ctx.world > 1 else "single GPU",
           "collective": ("none (no collective call): the gradient kernel stores every finished tile of rows into all "
                          "GPUs' copies of the [B,T] result -- " + gather_path(ps) + " -- + one flag barrier "
                          "(mbqc_peer_barrier), all inside the timed region" if ctx.world > 1 else "none"),
           "roofline": {"bound": "hbm", "achieved": per / ctx.world / (ms * 1e-3) / 1e9,
                        "frac": per / ctx.world / (ms * 1e-3) / 1e9 / peak,
                        "algorithmic_bytes_per_launch": per // ctx.world,
                        "note": "per GPU; 256 B per gradient (16 angles in, 16 derivatives out), 32 pattern evaluations each"},
           "l2": "angle matrix 128 MiB + gradients 128 MiB per pass (> L2)"}
    if ms_nccl is not None:
        res["nccl_all_gather"] = {"ms": ms_nccl, "value": B / (ms_nccl * 1e-3), "unit": "gradients/s",
                                  "what": "local kernel + ONE NCCL all_gather of the gradients (dist.gather_slices), same timed region",
                                  "max_abs_diff_vs_peer_store": agree}
    if gold:
        got = grad[0].cpu().numpy()
        res["parity"] = {"against": "tests/golden/gradients.json c4.psr (mentpy.gradients.get_gradient on the reference), row 0"
                                    + (" of the all-gathered result" if ctx.world > 1 else ""),
                         "max_abs_diff": float(np.abs(got - np.asarray(gold["psr"])).max())}
    assert bool(torch.isfinite(grad).all().item()) and tuple(grad.shape) == (B, T)
    return res


def _stream_run(ctx, w, fuse, group=None, reps=2):
    """One linear_cluster(w+16, window_size=w) pattern through the streaming backend; returns
    (seconds per pattern by CUDA events (max over ranks), output amplitudes, schedule)."""
    import mentpy_b200 as mb

    torch = ctx.torch
    gs = mb.templates.linear_cluster(w + 16)
    kw = {} if group is None else {"group": group}
    ps = mb.PatternSimulator(gs, backend="cuda-sv-stream", window_size=w, fuse=fuse, **kw)
    ang = np.random.default_rng(4).uniform(0, 2 * np.pi, w + 15)
    got = ps.run(ang)  # warm-up: allocation (+ CUDA IPC mapping when sharded)
    best = None
    stream = torch.cuda.current_stream(ctx.dev)
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if group is None:
            ctx.barrier()
        else:
            torch.cuda.synchronize(ctx.dev)
        e0.record(stream)
        got = ps.run(ang)
        e1.record(stream)
        torch.cuda.synchronize(ctx.dev)
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    if group is None:
        best = ctx.max_over_ranks([best])[0]
    sched = ps.simulator.last_schedule
    eng = getattr(ps.simulator, "_engine", None)
    if eng is not None:
        eng.release()
    del ps
    torch.cuda.empty_cache()
    return best * 1e-3, got, ang, sched


def config_c5(ctx, peak, fuse=5):
    """Large-window streaming state vector.  N = 1: w = 32 (64 GiB in place).  N > 1: sharded by the
    top log2 N qubits, shard-slot measurements read the partner's half over NVLink peer memory inside
    the pair-reduction kernel (stream_exchange_kernel); strong (w = 32) and weak (w = 32 + log2 N)."""
    from mentpy_b200.streaming import ExchangePass

    g = ctx.world.bit_length() - 1
    if (1 << g) != ctx.world:
        return {"skipped": "needs a power-of-two number of ranks"}
    res = {"pattern": "linear_cluster(w+16, window_size=w), one angle set (BASELINE configs[4])", "fuse": fuse,
           "unit": "patterns/s", "l2": "state >> L2"}
    cases = [("w32", 32)] if ctx.world == 1 else [("strong_w32", 32), (f"weak_w{32 + g}", 32 + g)]
    for tag, w in cases:
        s, got, ang, sched = _stream_run(ctx, w, fuse)
        want = linear_cluster_closed_form(ang)
        per_gpu = sched.algorithmic_bytes / ctx.world
        res[tag] = {"window": w, "state_GiB_total": 16 * 2.0**w / 2**30, "gpus": ctx.world,
                    "ms_per_pattern": s * 1e3, "value": 1.0 / s, "passes": len(sched.passes),
                    "exchange_passes": sum(isinstance(p, ExchangePass) for p in sched.passes),
                    "roofline": {"bound": "hbm", "achieved": per_gpu / s / 1e9, "frac": per_gpu / s / 1e9 / peak,
                                 "streamed_GBps_per_gpu": sched.streamed_bytes / ctx.world / s / 1e9,
                                 "frac_streamed": sched.streamed_bytes / ctx.world / s / 1e9 / peak,
                                 "note": "per GPU; achieved / frac use the ALGORITHMIC bytes 2*16*2^n per measurement at live window n "
                                         f"(SURVEY 8d); K = {fuse} measurements are fused per pass over the state, so frac > 1 is "
                                         "legitimate -- frac_streamed is what actually crosses HBM"},
                    "parity": {"against": "closed form J(-th_{L-2})...J(-th_0)|+> of the linear cluster",
                               "infidelity": infidelity(got, want)}}
    if ctx.world > 1:  # 1-GPU == sharded agreement at a window one GPU runs quickly
        w = 28
        _, got_sh, ang, _ = _stream_run(ctx, w, fuse, reps=1)
        if ctx.rank == 0:
            _, got_1, _, _ = _stream_run(ctx, w, fuse, group=False, reps=1)
            res["sharded_vs_single_gpu"] = {"window": w, "max_abs_diff": float(np.abs(got_sh - got_1).max()),
                                            "infidelity": infidelity(got_sh, got_1)}
        ctx.barrier()
    first = res[cases[0][0]]
    res["value"], res["ms"], res["roofline"] = first["value"], first["ms_per_pattern"], first["roofline"]
    return res


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import mentpy_b200 as mb
    from mentpy_b200 import _lib

    ctx = Ctx()
    torch, dist = ctx.torch, ctx.dist
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    K = args.steps
    W = max(args.warmup, 3)

    gs = mb.templates.grid_cluster(ROWS, COLS)
    T = len(gs.trainable_nodes)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    sim = ps.simulator
    k = len(gs.output_nodes)
    assert (T, k) == (N_ANGLES, N_OUT)

    # pool of batches: angles + outputs together exceed L2, every rank has its own angle stream
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED + 1000 * rank)
    angles = torch.rand((POOL, BATCH, T), generator=gen, device=dev, dtype=torch.float64) * (2 * np.pi)
    gold = golden_case(("grid_cluster", [2, 6]), "numpy-sv", "sv")
    if gold:
        angles[0, 0] = torch.tensor(gold["angles"], dtype=torch.float64)
    outs = torch.empty((POOL, BATCH, 2**k), dtype=torch.complex128, device=dev)
    status = torch.zeros(BATCH, dtype=torch.int32, device=dev)
    lib = _lib.load()
    plan = sim._full_plan()
    stream = torch.cuda.current_stream(dev)
    a_ptr = [angles[j].data_ptr() for j in range(POOL)]
    o_ptr = [outs[j].data_ptr() for j in range(POOL)]
    s_ptr = status.data_ptr()
    run_sv = lib.mbqc_run_batch_sv

    def step(i, cuda_stream):
        j = i % POOL
        rc = run_sv(plan.handle, a_ptr[j], T, None, _lib.INPUT_PLUS, BATCH, o_ptr[j], _lib.OUT_SV, s_ptr, cuda_stream)
        if rc != 0:
            _lib.check(rc)

    launches_before = _lib.launch_count()
    for i in range(W):
        step(i, stream.cuda_stream)
    ctx.barrier()
    launches_per_step = (_lib.launch_count() - launches_before) // W
    jit_info = lib.mbqc_jit_info().decode()
    jit_on = "compiled=0 from_disk=0" not in jit_info and "failures=0" in jit_info
    kernel_name = ("mbqc_jit_sv (mentpy_b200/csrc/sv_jit_src.inc: the pattern's kernel, specialised and compiled at plan time with NVRTC)"
                   if jit_on else "sv_lean_kernel<3,128,false,0> (mentpy_b200/csrc/sv_lean.cuh)")
    facts_key = "mbqc_jit_sv_c2" if jit_on else "sv_lean_kernel_c2"

    # the K-step block: launch-bound (one ~3 us kernel per step), so it is captured once as a CUDA
    # graph (independent batches in parallel branches) and replayed; --no-graph launches directly
    if args.no_graph:
        def block():
            for i in range(K):
                step(i, stream.cuda_stream)
        mode = "direct stream launches"
    else:
        g = graph_of(ctx, [(lambda s, i=i: step(i, s)) for i in range(K)], args.graph_branches)
        g.replay()
        ctx.barrier()
        block = g.replay
        mode = f"CUDA graph of the {K} steps in {min(args.graph_branches, K)} parallel branches, replayed"

    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    block_ms = timed_blocks(ctx, block, args.repeats)
    sampler.sample_once()
    sampler.stop()
    ms = float(np.median(block_ms))
    ms_per_rank = ctx.gather_per_rank(ms)
    assert int(status.max().item()) == 0, "kernel reported a bad norm"
    parity = None
    if gold:
        got = outs[0, 0].cpu().numpy()
        parity = {"against": "tests/golden/sim_cases.json (recorded from the reference), row 0 of batch 0",
                  "infidelity": infidelity(got, cplx(gold["output"])),
                  "max_abs_diff": float(np.abs(got - cplx(gold["output"])).max())}

    # the same K steps with plain stream launches (no graph), for the record
    def direct():
        for i in range(K):
            step(i, stream.cuda_stream)
    ms_nograph = float(np.median(timed_blocks(ctx, direct, 5)))

    fp64_peak = fp64_peak_tflops(ctx)

    # end-to-end through the public API with host buffers (pinned), every step: H2D + kernel + D2H
    e2e = None
    host_pool = 4
    if not args.skip_e2e:
        hgen = torch.Generator()
        hgen.manual_seed(SEED + 1000 * rank + 7)
        h_angles = [torch.empty((BATCH, T), dtype=torch.float64).pin_memory() for _ in range(host_pool)]
        for h in h_angles:
            h.uniform_(0.0, 2 * np.pi, generator=hgen)
        e2e_steps = max(3, min(K, 50))

        # (a) one blocking call per step
        for i in range(2):
            ps.run_batch(h_angles[i % host_pool], copy=False)
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            res = ps.run_batch(h_angles[i % host_pool], copy=False)
        torch.cuda.synchronize(dev)
        e2e_sync_s = ctx.max_over_ranks([time.perf_counter() - t0])[0]
        assert res.shape == (BATCH, 2**k)

        # (b) the asynchronous form of the same call, three steps in flight: the H2D copy of step
        # n+1 overlaps the kernels / result transfer of step n.  Every step's angles cross PCIe
        # and every step's amplitudes land in host memory (and are touched) inside the timed region.
        def pipelined(n_steps):
            pend, acc = [], 0.0
            for i in range(n_steps):
                pend.append(ps.run_batch_async(h_angles[i % host_pool]))
                if len(pend) == 3:
                    r = pend.pop(0).result()
                    acc += r[0, 0].real + r[-1, -1].real
            for h in pend:
                r = h.result()
                acc += r[0, 0].real + r[-1, -1].real
            return r, acc

        pipelined(5)
        e2e_runs = []
        for _ in range(5):
            ctx.barrier()
            t0 = time.perf_counter()
            res, _acc = pipelined(e2e_steps)
            torch.cuda.synchronize(dev)
            e2e_runs.append(time.perf_counter() - t0)
            assert res.shape == (BATCH, 2**k) and np.isfinite(_acc)
        e2e_runs = ctx.max_over_ranks(e2e_runs)
        e2e_s = float(np.median(e2e_runs))
        floor = host_copy_floor(ctx, BATCH * T * 8, BATCH * (2**k) * 16)
        floor_ms = max(floor["h2d_only_ms"], floor["d2h_only_ms"])  # full duplex: the slower direction bounds a step
        e2e = {"value": world * BATCH * e2e_steps / e2e_s, "unit": "evals/s",
               "h2d_bytes_per_step": BATCH * T * 8, "d2h_bytes_per_step": BATCH * (2**k) * 16,
               "steps": e2e_steps, "repeats": len(e2e_runs), "ms_per_step": 1e3 * e2e_s / e2e_steps,
               "host_floor_ms": floor_ms, "host_floor": floor,
               "host_floor_note": "the step's 5.24 MB H2D and 4.19 MB D2H as raw pinned DMA copies on the same "
                                  f"{world} rank(s) at once, no kernels (max over ranks): each direction alone and both together; "
                                  "host_floor_ms = the slower direction alone (PCIe is full duplex); aggregate over ranks at "
                                  f"both_directions_ms: {world * (BATCH * T * 8 + BATCH * (2**k) * 16) / (floor['both_directions_ms'] * 1e-3) / 1e9:.1f} GB/s",
               "api": "PatternSimulator(gs, backend='cuda-sv').run_batch_async(pinned host angles).result() -> host amplitudes, "
                      "three calls in flight (C ABI mbqc_run_batch_sv_host_submit / mbqc_host_wait: chunked H2D DMA + kernels storing "
                      "CTA-coalesced results straight into the mapped page-locked output buffer; consecutive calls on alternating stream sets)",
               "blocking_call": {"value": world * BATCH * e2e_steps / e2e_sync_s, "ms_per_step": 1e3 * e2e_sync_s / e2e_steps,
                                 "api": "run_batch(pinned host angles, copy=False), one blocking call per step"}}

    peak, peak_src = measured_peak_gbs()
    configs = None
    if not args.skip_configs:
        configs = {}
        legs = [("C4", config_c4), ("C5", config_c5)] if world > 1 else \
               [("C1", config_c1), ("C3", config_c3), ("C4", config_c4), ("C5", config_c5)]
        for name, fn in legs:
            if args.only_configs and name not in args.only_configs.split(","):
                continue
            try:
                configs[name] = fn(ctx, peak)
            except Exception as e:  # a failed leg must not lose the headline line
                configs[name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
            ctx.barrier()
        annotate_fp64(configs, fp64_peak, world)

    if rank == 0:
        ms_per_step = ms / K
        value = world * BATCH * K / (ms * 1e-3)
        achieved = BATCH * ALGO_BYTES_PER_EVAL / (ms_per_step * 1e-3) / 1e9
        facts = ncu_facts()
        fp64_inst = facts.get(facts_key + "_fp64_inst_per_eval")
        total_inst = facts.get(facts_key + "_inst_per_eval")
        evals_per_gpu = BATCH / (ms_per_step * 1e-3)
        fp64 = {"peak_measured_tflops": fp64_peak,
                "peak_how": "own DFMA microbenchmark (mbqc_probe_fp64_fma: 8 independent chains per thread, CUDA events, best of 5)"}
        binds = "fp64" if jit_on else "issue"
        binds_note = ("FP64 pipe: the specialised kernel is ~63 % FP64 instructions (393 of 621 per evaluation) and ncu shows the FP64 "
                      "pipe as the busiest unit (70.9 % at B = 2^20, DRAM 37 %); HBM traffic is I/O only -- profiles/README.md"
                      if jit_on else
                      "neither roof: the general kernels are bound by instruction issue (an FP64 warp instruction holds the dispatch "
                      "port 2 cycles: cycles per warp ~= non-FP64 + 2 x FP64 instructions); HBM traffic is I/O only -- profiles/README.md")
        if fp64_inst:
            # every FP64 warp instruction counted as one FMA (2 flops per lane): pipe occupancy, not useful flops
            fp64["achieved_tflops"] = evals_per_gpu * fp64_inst * 2 / 1e12
            fp64["frac"] = fp64["achieved_tflops"] / fp64_peak
            fp64["fp64_inst_per_eval"] = fp64_inst
            fp64["inst_per_eval"] = total_inst
            fp64["source"] = "instruction counts per evaluation from the committed ncu capture (profiles/traffic.json)"
        cores = os.cpu_count() or 1
        try:
            os.sched_setaffinity(0, range(cores))  # the CPU leg uses every host core again
        except Exception:
            pass
        cpu_rate, cpu_s = (None, 0.0) if args.skip_cpu else cpu_port_rate(evals_per_core=2048, cores=cores)
        cfg = bench_config(world)
        line = {
            "metric": "pattern_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": K, "warmup": W, "repeats": args.repeats, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "timing": {"how": f"the {K}-step block timed {args.repeats} times (CUDA events on the launching stream, barrier + "
                              "synchronize on both sides of every block, max over ranks per block); value = median block",
                       "block_ms_median": ms, "block_ms_min": float(np.min(block_ms)), "block_ms_max": float(np.max(block_ms)),
                       "first_block_ms": block_ms[0], "launch_mode": mode,
                       "value_first_block": world * BATCH * K / (block_ms[0] * 1e-3),
                       "value_stream_launch": world * BATCH * K / (ms_nograph * 1e-3)},
            "parity": parity,
            "e2e": e2e,
            "ms_per_rank": [round(v, 5) for v in ms_per_rank],
            "gpu_launches": int(K * launches_per_step),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": facts.get(facts_key + "_bytes_per_launch"),
                         "peak_source": peak_src, "kernel": kernel_name, "jit": jit_info,
                         "algorithmic_bytes_per_launch": BATCH * ALGO_BYTES_PER_EVAL,
                         "fp64": fp64, "binds": binds, "binds_note": binds_note},
            "cpu_baseline": {"value": cpu_rate, "unit": "evals/s", "cores": cores, "kind": "port",
                             "sample": f"{2048 * cores} angle sets of the same workload, one single-threaded process per core, "
                                       f"{cpu_s:.1f} s (oracle/dense_port.py: reference algorithm incl. dense kron operators)"},
            "clocks": sampler.summary(),
            "configs": configs,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--repeats", type=int, default=50, help="how many times the K-step block is timed (median reported)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--graph-branches", type=int, default=GRAPH_BRANCHES, help="parallel branches of the CUDA graph (independent batches overlap)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step directly instead of replaying a CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: no cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: no host end-to-end leg")
    ap.add_argument("--skip-configs", action="store_true", help="profiling runs: headline only")
    ap.add_argument("--only-configs", default="", help="comma list out of C1,C3,C4,C5")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
