"""Run under torchrun: batch-split run_batch / psr gradient across GPUs == single-GPU result."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import mentpy_b200 as mb
from mentpy_b200.dist import psr_gradient_distributed, run_batch_distributed
from mentpy_b200.gradients import psr_gradient_batched


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
    g = psr_gradient_distributed(ps, ang[:515], tgt)
    gref = psr_gradient_batched(ps, torch.from_numpy(ang[:515]).cuda(), tgt)
    ok = ok and torch.equal(g, gref)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print("BATCH_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
