"""Under torchrun (N GPUs): where the time of the replicated-result gradient step goes.  BASELINE
config 4 (2^20 vectors, grid_cluster(4,5)) split over the ranks; CUDA events, max over ranks:
  local      plain kernel, rows stored locally only (mbqc_psr_grad_batch)
  push1      replicated-result kernel with ONE destination (its own copy): cost of the staged read-out
  pushN      the same kernel storing into all N copies (NVLink peer stores)
  barrier    mbqc_peer_barrier alone
  step       pushN + barrier (what dist.psr_gradient_distributed runs), Python call path included
  nccl       local + all_gather_into_tensor
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import mentpy_b200 as mb
from mentpy_b200 import _lib
from mentpy_b200.dist import ReplicatedResult, psr_gradient_distributed, slice_bounds
from mentpy_b200.gradients import psr_gradient_batched


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B = 1 << 20
    gs = mb.templates.grid_cluster(4, 5)
    T = len(gs.trainable_nodes)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    sim = ps.simulator
    gen = torch.Generator(device=dev)
    gen.manual_seed(4)
    full = torch.rand((B, T), generator=gen, device=dev, dtype=torch.float64) * (2 * np.pi)
    tgt = torch.full((16,), 0.25, dtype=torch.complex128, device=dev)
    lo, hi = slice_bounds(B, rank, world)
    part = full[lo:hi]
    n = hi - lo
    res = ReplicatedResult(B, T)
    plan = sim._full_plan()
    status = torch.empty(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def push(n_dst):
        dst, cnt = res.destinations()
        _lib.check(lib.mbqc_psr_grad_batch_push(plan.handle, part.data_ptr(), T, None, 0, n, tgt.data_ptr(), C.c_double(1.5),
                                                dst, min(cnt, n_dst), lo, None, status.data_ptr(), stream))

    def push_mc():
        _lib.check(lib.mbqc_psr_grad_batch_multicast(plan.handle, part.data_ptr(), T, None, 0, n, tgt.data_ptr(), C.c_double(1.5),
                                                     C.c_void_p(res.mc_ptr + res.copy * res.nbytes), lo, None, status.data_ptr(), stream))

    def timed(fn, reps=10, rounds=5):
        out = []
        for _ in range(rounds):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / reps)
        t = torch.tensor([float(np.median(out))], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rows = {}
    for _ in range(2):
        psr_gradient_batched(ps, part, tgt); push(1); push(world); res.barrier(); psr_gradient_distributed(ps, full, tgt)
    rows["local"] = timed(lambda: psr_gradient_batched(ps, part, tgt))
    rows["push1"] = timed(lambda: push(1))
    rows["pushN"] = timed(lambda: push(world))
    if res.mc_ptr:
        push_mc()
        res.barrier()
        torch.cuda.synchronize()
        ref = psr_gradient_batched(ps, full[:4096], tgt)
        rows["multicast_max_abs_diff_rows_0_4095"] = float((res.tensor[:4096] - ref).abs().max().item()) if rank == 0 else 0.0
        rows["pushMC"] = timed(push_mc)
        rows["pushMC+barrier"] = timed(lambda: (push_mc(), res.barrier()))
    rows["barrier"] = timed(lambda: res.barrier(), reps=50)
    rows["pushN+barrier"] = timed(lambda: (push(world), res.barrier()))
    rows["step"] = timed(lambda: psr_gradient_distributed(ps, full, tgt))
    rows["nccl"] = timed(lambda: psr_gradient_distributed(ps, full, tgt, fused=False))
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        psr_gradient_distributed(ps, full, tgt)
    rows["host_enqueue_ms_per_step"] = (time.perf_counter() - t0) * 1e3 / 50
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"gpus": world, "rows_per_gpu": n, "cta": os.environ.get("MBQC_GRAD_CTA", "128"), "multicast": bool(res.mc_ptr),
                          "ms": {k: round(v, 4) for k, v in rows.items()}}), flush=True)
    dist.barrier()
    res.release()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
