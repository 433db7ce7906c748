#!/usr/bin/env python
"""Benchmark of the MBQC pattern-evaluation hot path (BASELINE.json metric: pattern evals/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (N = 1 and per rank for N > 1, weak scaling): BASELINE.json configs[1] --
grid_cluster(2,6) state-vector pattern, 65,536 random angle sets per step, outputs [B,4]
complex128.  A step = one pass of the hot path over one batch = ONE kernel launch.  Batches rotate
through a pool whose angles+outputs exceed the 126 MB L2, so no step finds its inputs cached.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over
ranks); `e2e` = the same metric through PatternSimulator.run_batch with HOST buffers (pinned
H2D + D2H inside the timed region); `roofline` = algorithmic bytes / step time vs measured HBM
peak; `cpu_baseline` = the dense numpy port of the reference (oracle/dense_port.py) on host cores.
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
SEED = 1
ALGO_BYTES_PER_EVAL = 8 * 10 + 16 * 4  # 8*T angles in + 16*2^k amplitudes out = 144 B
L2_BYTES = 126 * 1024 * 1024


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("sv_reg_kernel_c2_bytes_per_launch")
        except Exception:
            return None
    return None


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


def cpu_port_rate(evals_per_core, cores, repeats=1, pool=None):
    """evals/s of the dense reference port using `cores` processes (each single-threaded)."""
    import multiprocessing as mp

    pat_json = _pattern_json()
    rng = np.random.default_rng(SEED)
    chunks = [rng.uniform(0, 2 * np.pi, (evals_per_core, 10)) for _ in range(cores)]
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None
    best = None
    try:
        if pool is not None:
            pool.map(_cpu_worker, [(pat_json, c[:2]) for c in chunks])  # warm the workers
        for _ in range(repeats):
            t0 = time.perf_counter()
            if pool is not None:
                pool.map(_cpu_worker, [(pat_json, c) for c in chunks])
            else:
                _cpu_worker((pat_json, chunks[0]))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    finally:
        if own and pool is not None:
            pool.close()
            pool.join()
    return evals_per_core * cores / best, best


def run_reference_arm(args):
    """--impl reference: the reference algorithm's CPU port on all host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    os.environ["OMP_NUM_THREADS"] = "1"
    cores = os.cpu_count() or 1
    per_core = 48  # ~0.3 s of work per core per step
    pat_json = _pattern_json()
    rng = np.random.default_rng(SEED)
    pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None
    sample = per_core * cores

    def step():
        chunks = [(pat_json, rng.uniform(0, 2 * np.pi, (per_core, 10))) for _ in range(cores)]
        if pool is not None:
            pool.map(_cpu_worker, chunks)
        else:
            _cpu_worker(chunks[0])

    steps = min(args.steps, 40)
    for _ in range(min(args.warmup, 3)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    if pool is not None:
        pool.close()
        pool.join()
    value = sample * steps / dt
    line = {
        "impl": "reference", "metric": "pattern_evals_per_s", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "grid_cluster(2,6) statevector, random angle sets (BASELINE configs[1])",
                   "pattern": "grid_cluster(2,6)", "backend": "numpy-sv algorithm (dense kron operators)",
                   "evals_per_step": sample, "window": 3, "measurements": 10},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} angle sets per step x {steps} steps, one single-threaded process per core "
                                   "(oracle/dense_port.py, restates np_simulator_sv.py incl. dense kron operators)"},
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
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import mentpy_b200 as mb
    from mentpy_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pin_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    gs = mb.templates.grid_cluster(ROWS, COLS)
    T = len(gs.trainable_nodes)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    sim = ps.simulator
    k = len(gs.output_nodes)

    # pool of batches: angles + outputs together exceed L2, every rank has its own angle stream
    per_batch = BATCH * (8 * T + 16 * 2**k)
    pool = max(4, int(np.ceil(1.5 * L2_BYTES / per_batch)))
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED + 1000 * rank)
    angles = torch.rand((pool, BATCH, T), generator=gen, device=dev, dtype=torch.float64) * (2 * np.pi)
    outs = torch.empty((pool, BATCH, 2**k), dtype=torch.complex128, device=dev)
    status = torch.empty(BATCH, dtype=torch.int32, device=dev)
    lib = _lib.load()
    plan = sim._full_plan()
    stream = torch.cuda.current_stream(dev)
    a_ptr = [angles[j].data_ptr() for j in range(pool)]
    o_ptr = [outs[j].data_ptr() for j in range(pool)]
    s_ptr = status.data_ptr()
    run_sv = lib.mbqc_run_batch_sv

    def step(i, cuda_stream):
        j = i % pool
        rc = run_sv(plan.handle, a_ptr[j], T, None, _lib.INPUT_PLUS, BATCH, o_ptr[j], _lib.OUT_SV, s_ptr, cuda_stream)
        if rc != 0:
            _lib.check(rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        step(i, stream.cuda_stream)
    barrier()
    gathered = torch.empty((world, BATCH, 2**k), dtype=torch.complex128, device=dev) if world > 1 else None

    # the step loop is launch-bound (one ~few-us kernel per step): capture one pass over the pool
    # in a CUDA graph and replay it; leftover steps are launched directly
    graph = graph_tail = None
    n_tail = args.steps % pool

    def capture(first, count):
        g = torch.cuda.CUDAGraph()
        cap_stream = torch.cuda.Stream(device=dev)
        side = [torch.cuda.Stream(device=dev) for _ in range(args.graph_branches - 1)]
        with torch.cuda.graph(g, stream=cap_stream):
            cur = torch.cuda.current_stream(dev)
            for sd in side:
                sd.wait_stream(cur)
            for j in range(first, first + count):  # independent batches: parallel branches, joined at the end
                br = j % args.graph_branches
                step(j, (cur if br == 0 else side[br - 1]).cuda_stream)
            for sd in side:
                cur.wait_stream(sd)
        g.replay()
        return g

    if not args.no_graph:
        graph = capture(0, pool)
        if n_tail:
            graph_tail = capture(0, n_tail)  # the K mod pool leftover steps, same launch mode
        barrier()

    def run_steps(n):
        done = 0
        if graph is not None:
            while n - done >= pool:
                graph.replay()
                done += pool
            if graph_tail is not None and n - done == n_tail:
                graph_tail.replay()
                done += n_tail
        for i in range(done, n):
            step(i, stream.cuda_stream)

    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record(stream)
    run_steps(args.steps)
    e1.record(stream)
    sampler.sample_once()
    barrier()
    sampler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:  # outside the timed region: the path has no exchange step; sanity gather
        dist.all_gather_into_tensor(gathered.view(-1), outs[(args.steps - 1) % pool].view(-1))
        torch.cuda.synchronize(dev)
        assert bool(torch.isfinite(gathered.real).all().item())
    launches = args.steps  # one kernel per step (graph replays execute `pool` kernel nodes each)
    ms_per_rank = [ms]
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = ms
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_per_rank = [float(v) for v in t.tolist()]
        ms = max(ms_per_rank)  # the job is as slow as its slowest rank
    assert int(status.max().item()) == 0, "kernel reported a bad norm"

    # same K steps with plain stream launches (no graph), for the record
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for i in range(args.steps):
        step(i, stream.cuda_stream)
    e3.record(stream)
    barrier()
    ms_nograph = e2.elapsed_time(e3)

    # end-to-end through the public API with host buffers (pinned), every step: H2D + kernel + D2H
    e2e = None
    host_pool = min(pool, 4)
    if not args.skip_e2e:
        # host-side inputs are produced on the host, directly in page-locked memory
        hgen = torch.Generator()
        hgen.manual_seed(SEED + 1000 * rank + 7)
        h_angles = [torch.empty((BATCH, T), dtype=torch.float64).pin_memory() for _ in range(host_pool)]
        for h in h_angles:
            h.uniform_(0.0, 2 * np.pi, generator=hgen)
        e2e_steps = max(3, min(args.steps, 50))

        def max_over_ranks(sec):
            if world > 1:
                t = torch.tensor([sec], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())
            return sec

        # (a) one blocking call per step
        for i in range(2):
            ps.run_batch(h_angles[i % host_pool], copy=False)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            res = ps.run_batch(h_angles[i % host_pool], copy=False)
        torch.cuda.synchronize(dev)
        e2e_sync_s = max_over_ranks(time.perf_counter() - t0)
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
        barrier()
        t0 = time.perf_counter()
        res, _acc = pipelined(e2e_steps)
        torch.cuda.synchronize(dev)
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        assert res.shape == (BATCH, 2**k) and np.isfinite(_acc)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ms_per_step = ms / args.steps
        value = world * BATCH * args.steps / (ms * 1e-3)
        achieved = BATCH * ALGO_BYTES_PER_EVAL / (ms_per_step * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        try:
            os.sched_setaffinity(0, range(cores))  # the CPU leg uses every host core again
        except Exception:
            pass
        cpu_rate, cpu_s = (None, 0.0) if args.skip_cpu else cpu_port_rate(evals_per_core=4096, cores=cores)
        line = {
            "metric": "pattern_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "grid_cluster(2,6) statevector, 65,536 random angle sets per step per GPU (BASELINE configs[1])",
                       "pattern": "grid_cluster(2,6)", "backend": "cuda-sv", "batch_per_gpu": BATCH,
                       "window": 3, "measurements": 10, "output": "sv [B,4] complex128",
                       "launch_mode": "direct stream launches" if graph is None else f"CUDA graph of {pool} steps in {args.graph_branches} parallel branches, replayed",
                       "l2": f"inputs rotate through a pool of {pool} batches ({pool * per_batch / 2**20:.0f} MiB > 126 MiB L2)",
                       "parallelism": f"batch-split x{world}, no collective on the data path" + (" (outputs of the last step are all-gathered once after the timed region as a check)" if world > 1 else "")},
            "e2e": None if args.skip_e2e else {
                "value": world * BATCH * e2e_steps / e2e_s, "unit": "evals/s",
                "h2d_bytes_per_step": BATCH * T * 8, "d2h_bytes_per_step": BATCH * (2**k) * 16,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "api": "PatternSimulator(gs, backend='cuda-sv').run_batch_async(pinned host angles).result() -> host amplitudes, "
                       "three calls in flight (C ABI mbqc_run_batch_sv_host_submit / mbqc_host_wait: chunked H2D DMA + kernels storing "
                       "CTA-coalesced results straight into the mapped page-locked output buffer; consecutive calls on alternating stream sets)",
                "blocking_call": {"value": world * BATCH * e2e_steps / e2e_sync_s, "ms_per_step": 1e3 * e2e_sync_s / e2e_steps,
                                  "api": "run_batch(pinned host angles, copy=False), one blocking call per step"}},
            "ms_per_rank": [round(v, 4) for v in ms_per_rank],
            "value_stream_launch": world * BATCH * args.steps / (ms_nograph * 1e-3),
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                         "peak_source": peak_src, "kernel": "sv_reg_kernel<3,false>",
                         "algorithmic_bytes_per_launch": BATCH * ALGO_BYTES_PER_EVAL,
                         "note": "register-resident batched regime is FP64-pipe/launch bound, not HBM bound (SURVEY 8d); see DESIGN.md"},
            "cpu_baseline": {"value": cpu_rate, "unit": "evals/s", "cores": cores, "kind": "port",
                             "sample": f"{4096 * cores} angle sets of the same workload, one single-threaded process per core, "
                                       f"{cpu_s:.1f} s (oracle/dense_port.py: reference algorithm incl. dense kron operators)"},
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--graph-branches", type=int, default=4, help="parallel branches of the CUDA graph (independent batches overlap)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step directly instead of replaying a CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: no cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: no host end-to-end leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
