"""Run under torchrun: batch-split run_batch / psr gradient across GPUs == single-GPU result."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import mentpy_b200 as mb
from mentpy_b200.dist import (psr_gradient_dataset_distributed, psr_gradient_distributed, run_batch_distributed,
                              sample_batch_distributed)
from mentpy_b200.gradients import psr_gradient_batched, psr_gradient_dataset


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    gs = mb.templates.grid_cluster(4, 5)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    B, T = 4099, len(gs.trainable_nodes)  # ragged split
    ang = np.random.default_rng(3).uniform(0, 2 * np.pi, (B, T))
    full = run_batch_distributed(ps, ang)
    ref = ps.run_batch(torch.from_numpy(ang).cuda())
    ok = torch.equal(full, ref)
    tgt = np.full(16, 0.25)
    gref = psr_gradient_batched(ps, torch.from_numpy(ang[:515]).cuda(), tgt)
    g = psr_gradient_distributed(ps, ang[:515], tgt, fused=False)  # local kernel + NCCL all_gather
    ok = ok and torch.equal(g, gref)
    # replicated result over NVLink peer memory: general kernels + peer copies (small batch), then
    # the specialised kernel storing into every GPU's copy itself; repeated calls alternate copies
    from mentpy_b200 import _lib
    lib = _lib.load()
    for mode in (0, 2, 2, 2):
        prev = lib.mbqc_jit_set_mode(mode)
        g = psr_gradient_distributed(ps, ang[:515], tgt)
        ok = ok and bool((g - gref).abs().max() < 1e-12)
        lib.mbqc_jit_set_mode(prev)
    gbig_ref = psr_gradient_batched(ps, torch.from_numpy(ang).cuda(), tgt)
    lib.mbqc_jit_set_mode(2)
    for _ in range(3):
        gbig = psr_gradient_distributed(ps, ang, tgt)  # ragged split: 2050 + 2049 rows
        ok = ok and bool((gbig - gbig_ref).abs().max() < 1e-12)
    # the same through CUDA-IPC mapped peer copies (what runs where no multicast mapping exists)
    from mentpy_b200.dist import ReplicatedResult
    ipc = ReplicatedResult(B, T, multicast=False)
    ok = ok and ipc.mc_ptr == 0
    for _ in range(3):
        gipc = psr_gradient_distributed(ps, ang, tgt, result=ipc)
        ok = ok and bool((gipc - gbig_ref).abs().max() < 1e-12)
    lib.mbqc_jit_set_mode(0)  # general kernels + peer copies
    gipc = psr_gradient_distributed(ps, ang, tgt, result=ipc)
    ok = ok and bool((gipc - gbig_ref).abs().max() < 1e-12)
    torch.cuda.synchronize()
    dist.barrier()
    ipc.release()
    lib.mbqc_jit_set_mode(1)
    ok = ok and "failures=0" in lib.mbqc_jit_info().decode()
    # sampled shots: the Philox stream is indexed by the global shot number -> identical records
    pss = mb.PatternSimulator(gs, backend="cuda-sv", force0=False, seed=9)
    shots = sample_batch_distributed(pss, ang, seed=9, sample_offset=100)
    one = pss.sample_batch(torch.from_numpy(ang).cuda(), seed=9, sample_offset=100)
    ok = ok and torch.equal(shots.outcomes, one.outcomes) and torch.equal(shots.states, one.states)
    ok = ok and torch.equal(shots.x, one.x) and torch.equal(shots.prob, one.prob)
    # data-set averaged gradient, data items split across the ranks, one all_reduce
    rng = np.random.default_rng(5)
    S = 301
    ins = rng.normal(size=(S, 16)) + 1j * rng.normal(size=(S, 16)); ins /= np.linalg.norm(ins, axis=1, keepdims=True)
    tg = rng.normal(size=(S, 16)) + 1j * rng.normal(size=(S, 16)); tg /= np.linalg.norm(tg, axis=1, keepdims=True)
    gd, cd = psr_gradient_dataset_distributed(ps, ang[:7], tg, ins, return_cost=True)
    gd1, cd1 = psr_gradient_dataset(ps, torch.from_numpy(ang[:7]).cuda(), tg, ins, return_cost=True)
    ok = ok and torch.allclose(gd, gd1, atol=1e-13, rtol=0) and torch.allclose(cd, cd1, atol=1e-13, rtol=0)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print("BATCH_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
