"""CPU: host-side scheduling of the streaming / sharded regime (mentpy_b200/streaming.py) against
the oracles, with the kernels' index arithmetic emulated in numpy (tests/stream_numpy_engine.py).
world_size 2 and 4 run as real multi-process gloo jobs."""
import os
import sys

import numpy as np
import pytest

import mentpy_b200 as mb
from mentpy_b200.plan import lower
from mentpy_b200.streaming import ExchangePass, LocalPass, StreamExecutor, build_schedule
from oracle import matrix_free
from oracle.pattern_data import PatternData
from stream_numpy_engine import NumpyStreamEngine

from conftest import infidelity_pure

CASES = [("linear_cluster", [14], 6), ("linear_cluster", [20], 9), ("grid_cluster", [3, 6], 7),
         ("grid_cluster", [2, 8], 5), ("grid_cluster", [4, 4], 6), ("many_wires", [[4, 5, 3]], 6)]


def _want(gs, w, ang, inp=None):
    return matrix_free.run_sv_batch(PatternData.from_circuit(gs), ang, inp, window_size=w)[0]


@pytest.mark.parametrize("name,args,w", CASES)
@pytest.mark.parametrize("fuse", [1, 3, 5])
def test_single_rank_schedule_matches_oracle(name, args, w, fuse):
    gs = getattr(mb.templates, name)(*args)
    plan = lower(gs, window_size=w)
    ang = np.random.default_rng(3).uniform(0, 2 * np.pi, plan.n_angles)
    got = StreamExecutor(plan, NumpyStreamEngine(), 0, 0, fuse).run(ang)
    want = _want(gs, w, ang)
    assert infidelity_pure(got, want) < 1e-12
    assert np.allclose(got, want, atol=1e-9)  # incl. the reference's global phase


def test_schedule_accounting_and_fusion():
    gs = mb.templates.linear_cluster(26)
    plan = lower(gs, window_size=10)
    ang = np.zeros(plan.n_angles)
    s1 = build_schedule(plan, ang, 0, fuse=1)
    s4 = build_schedule(plan, ang, 0, fuse=4)
    assert len(s1.passes) == len(plan.steps) and len(s4.passes) < len(s1.passes) / 2
    assert s1.algorithmic_bytes == s4.algorithmic_bytes
    # 2*16*2^n per measurement at live window n: 16 full steps + halving tail
    full = 2 * 16 * 2**10
    # 25 measurements: 16 with an append at live window 10, then the tail reads 2^10, 2^9, ..., 2^2
    assert s1.algorithmic_bytes == 16 * full + sum(2 * 16 * 2**n for n in range(10, 1, -1))
    assert s4.streamed_bytes < 0.5 * s1.streamed_bytes
    assert all(isinstance(p, LocalPass) for p in s4.passes)
    s2 = build_schedule(plan, ang, shard_bits=2, fuse=4)
    assert any(isinstance(p, ExchangePass) for p in s2.passes)
    with pytest.raises(ValueError):
        build_schedule(plan, ang[:-1])
    with pytest.raises(ValueError):
        build_schedule(lower(mb.templates.linear_cluster(5)), np.zeros(4), shard_bits=1)


def test_haar_input_and_fixed_angles():
    from scipy.stats import unitary_group

    gs = mb.templates.grid_cluster(3, 5)
    gs[3] = mb.Ment("X")
    gs[7] = mb.Ment(0.4, "XY")
    plan = lower(gs, window_size=8)
    ang = np.random.default_rng(5).uniform(0, 2 * np.pi, plan.n_angles)
    inp = unitary_group.rvs(8, random_state=2)[:, 0]
    got = StreamExecutor(plan, NumpyStreamEngine(), 0, 0, 4).run(ang, inp)
    assert np.allclose(got, _want(gs, 8, ang, inp), atol=1e-9)


def _gloo_worker(rank, world, port, spec, w, fuse, seed, q, order="msb"):
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from stream_numpy_engine import NumpyStreamEngine as Eng

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        name, args = spec
        gs = getattr(mb.templates, name)(*args)
        plan = lower(gs, window_size=w, slot_order=order, shard_bits=world.bit_length() - 1)
        ang = np.random.default_rng(seed).uniform(0, 2 * np.pi, plan.n_angles)
        g = world.bit_length() - 1
        out = StreamExecutor(plan, Eng(dist), rank, g, fuse).run(ang)
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,spec,w,fuse,order", [(2, ("linear_cluster", [18]), 8, 3, "msb"),
                                                     (2, ("grid_cluster", [3, 6]), 7, 4, "lsb"),
                                                     (4, ("linear_cluster", [16]), 7, 2, "lsb"),
                                                     (4, ("grid_cluster", [2, 8]), 6, 5, "msb"),
                                                     (2, ("linear_cluster", [40]), 8, 5, "lsb"),
                                                     (4, ("linear_cluster", [24]), 9, 5, "shard-last"),
                                                     (2, ("grid_cluster", [3, 7]), 8, 4, "shard-last")])
def test_sharded_schedule_gloo(world, spec, w, fuse, order):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 7 + w + fuse * 13
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, spec, w, fuse, 11, q, order)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    name, args = spec
    gs = getattr(mb.templates, name)(*args)
    ang = np.random.default_rng(11).uniform(0, 2 * np.pi, len(gs.trainable_nodes))
    want = _want(gs, w, ang)
    assert infidelity_pure(got, want) < 1e-12
    assert np.allclose(got, want, atol=1e-9)
