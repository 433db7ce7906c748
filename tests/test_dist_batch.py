"""CPU (gloo, world_size 2 and 3): batch-split helpers of mentpy_b200/dist.py -- balanced contiguous
slices and the ragged all_gather -- plus the GPU torchrun check of run_batch_distributed."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from mentpy_b200.dist import slice_bounds


def test_slice_bounds_cover_and_balance():
    for total in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [slice_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slice_bounds(4, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from mentpy_b200.dist import gather_slices, slice_bounds as sb

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 11
        full = torch.arange(total * 4, dtype=torch.float64).reshape(total, 4)
        cfull = torch.complex(full, -full)
        lo, hi = sb(total, rank, world)
        got = gather_slices(full[lo:hi].clone(), total)
        gotc = gather_slices(cfull[lo:hi].clone(), total)
        ok = torch.equal(got, full) and torch.equal(gotc, cfull)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ragged_gather_gloo(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_run_batch_distributed_under_torchrun():
    script = os.path.join(ROOT, "tests", "multi_gpu_batch_check.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29641", script]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "BATCH_CHECK_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
