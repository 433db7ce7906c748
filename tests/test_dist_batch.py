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


class _FakeSim:
    """CPU stand-in with the two entry points the distributed helpers call: deterministic
    functions of (angles, data items, global shot index), so the gathered / reduced result has a
    closed form."""

    input_state = None

    def _dev(self):
        return torch.device("cpu")

    def sample_batch(self, angles, input_states=None, seed=None, sample_offset=0, **kw):
        from mentpy_b200.simulators.cuda_backends import SampledBatch

        n = angles.shape[0]
        shot = torch.arange(sample_offset, sample_offset + n, dtype=torch.float64)
        states = torch.complex(angles.sum(dim=1, keepdim=True) + shot[:, None], shot[:, None] * 0 + (seed or 0))
        oc = (shot.to(torch.int64)[:, None] + torch.arange(3)[None]).remainder(2).to(torch.int8)
        return SampledBatch(states, oc, oc[:, :1], oc[:, 1:2], shot / 100.0)


def _worker_dist_helpers(rank, world, port, q):
    import torch.distributed as dist

    import mentpy_b200.gradients as grads
    from mentpy_b200 import dist as mdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sim = _FakeSim()
        B, T = 13, 3
        ang = np.arange(B * T, dtype=np.float64).reshape(B, T)
        got = mdist.sample_batch_distributed(sim, ang, seed=7, sample_offset=40)
        want = sim.sample_batch(torch.from_numpy(ang), seed=7, sample_offset=40)
        ok = all(torch.equal(a, b) for a, b in zip(got, want))

        # data-set gradient: per-item gradient g_s(x) = x * (s + 1), cost_s = s -> means over ALL items
        def fake_dataset(sim_, a, targets, inputs, shift=1.5, return_cost=False):
            w = torch.as_tensor(np.asarray(targets)[:, 0].real, dtype=torch.float64)  # item ids travel in the targets
            g = a[..., None, :] * (w[:, None] + 1.0)
            return g.mean(dim=-2), (w.mean().expand(a.shape[:-1]) if a.dim() == 2 else w.mean())

        grads.psr_gradient_dataset = fake_dataset
        S = 7
        targets = np.arange(S, dtype=np.float64)[:, None] * np.ones((1, 2)) + 0j
        X = np.arange(8, dtype=np.float64).reshape(2, 4) + 1
        g, c = mdist.psr_gradient_dataset_distributed(sim, X, targets, None, return_cost=True)
        wmean = (np.arange(S) + 1).mean()
        ok = ok and np.allclose(g.numpy(), X * wmean) and np.allclose(c.numpy(), np.arange(S).mean())
        g1 = mdist.psr_gradient_dataset_distributed(sim, X[0], targets, None)
        ok = ok and np.allclose(g1.numpy(), X[0] * wmean)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_sampling_and_dataset_gradient_gloo(world):
    """Host logic of sample_batch_distributed (global shot offsets, ragged gather of every field)
    and psr_gradient_dataset_distributed (item split, size-weighted all_reduce), world 2 and 3."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 1000) + world
    procs = [ctx.Process(target=_worker_dist_helpers, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def _worker_batch_split(rank, world, port, q):
    import torch.distributed as dist

    import mentpy_b200.gradients as grads
    from mentpy_b200 import dist as mdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Sim(_FakeSim):
            def run_batch(self, angles, **kw):  # row b -> [sum of its angles, first angle]
                return torch.stack([angles.sum(dim=1), angles[:, 0]], dim=1).to(torch.complex128)

        sim = Sim()
        B, T = 11, 3  # ragged over 2 and 3 ranks
        ang = np.arange(B * T, dtype=np.float64).reshape(B, T)
        full = mdist.run_batch_distributed(sim, ang)
        want = sim.run_batch(torch.from_numpy(ang))
        ok = torch.equal(full, want)
        part, (lo, hi) = mdist.run_batch_distributed(sim, ang, gather=False)
        ok = ok and (lo, hi) == mdist.slice_bounds(B, rank, world) and torch.equal(part, want[lo:hi])
        # gradient split, NCCL / gloo all_gather form: g[b, i] = x[b, i] * (i + 1)
        grads.psr_gradient_batched = lambda s, a, target, shift=1.5: a * torch.arange(1, a.shape[1] + 1, dtype=a.dtype)
        g = mdist.psr_gradient_distributed(sim, ang, np.zeros(2), fused=False)
        ok = ok and torch.equal(g, torch.from_numpy(ang) * torch.arange(1, T + 1, dtype=torch.float64))
        gl, (lo2, hi2) = mdist.psr_gradient_distributed(sim, ang, np.zeros(2), gather=False)
        ok = ok and (lo2, hi2) == (lo, hi) and torch.equal(gl, g[lo:hi])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_batch_split_and_gradient_split_gloo(world):
    """Host logic of run_batch_distributed / psr_gradient_distributed (contiguous ragged slices, the
    all_gather form of the final gather, gather=False) on CPU, world 2 and 3.  The peer-memory form
    of the gather needs GPUs: tests/multi_gpu_batch_check.py under torchrun."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 1000) + world
    procs = [ctx.Process(target=_worker_batch_split, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
