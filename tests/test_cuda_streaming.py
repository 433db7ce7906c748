"""GPU: streaming regime (one large window in HBM) through the C ABI -- parity with the numpy
oracle at windows it can hold, the analytic linear-cluster oracle beyond, fused == unfused, and
(when >= 2 GPUs are visible) the NVLink-sharded run under torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import mentpy_b200 as mb
from conftest import ROOT, infidelity_pure
from oracle import matrix_free
from oracle.pattern_data import PatternData

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,args,w", [("linear_cluster", [20], 10), ("grid_cluster", [3, 7], 8),
                                         ("grid_cluster", [4, 5], 9), ("grid_cluster", [2, 12], 12),
                                         ("many_wires", [[4, 6, 5]], 7), ("linear_cluster", [36], 20)])
@pytest.mark.parametrize("fuse", [1, 4, 5])
def test_stream_matches_oracle(name, args, w, fuse):
    gs = getattr(mb.templates, name)(*args)
    ang = np.random.default_rng(9).uniform(0, 2 * np.pi, len(gs.trainable_nodes))
    ps = mb.PatternSimulator(gs, backend="cuda-sv-stream", window_size=w, fuse=fuse)
    got = ps.run(ang, output_form="sv")
    want = matrix_free.run_sv_batch(PatternData.from_circuit(gs), ang, window_size=w)[0]
    assert infidelity_pure(got, want) < 1e-10
    assert np.allclose(got, want, atol=1e-9)
    assert ps.outcomes == {v: 0 for v in ps.schedule_measure}
    # the batched register/smem kernels agree with the streaming kernels where both apply
    if w <= 12:
        batched = mb.PatternSimulator(gs, backend="cuda-sv", window_size=w).run_batch(ang[None, :])[0]
        assert np.allclose(got, batched, atol=1e-10)


def test_stream_input_state_and_dm_form():
    from scipy.stats import unitary_group

    gs = mb.templates.grid_cluster(3, 6)
    gs[2] = mb.Ment("X")
    ang = np.random.default_rng(1).uniform(0, 2 * np.pi, len(gs.trainable_nodes))
    inp = unitary_group.rvs(8, random_state=3)[:, 0]
    ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv-stream", window_size=10)
    rho = ps.run(ang, output_form="dm")
    want = matrix_free.run_sv_batch(PatternData.from_circuit(gs), ang, inp, window_size=10, output_form="dm")[0]
    assert np.abs(rho - want).max() < 1e-10


def test_stream_large_window_analytic_and_teleportation():
    """w = 27 (2 GiB state): beyond anything the reference can run; analytic oracle
    J(-th_{L-2})...J(-th_0)|in> and the teleportation identity at all-zero angles."""
    w = 27
    gs = mb.templates.linear_cluster(w + 16)   # odd length: identity channel at zero angles
    ps = mb.PatternSimulator(gs, backend="cuda-sv-stream", window_size=w, fuse=4)
    ang = np.random.default_rng(4).uniform(0, 2 * np.pi, w + 15)
    got = ps.run(ang, output_form="sv")
    assert infidelity_pure(got, matrix_free.linear_cluster_analytic(ang)[0]) < 1e-10
    st = np.array([0.6, 0.8j])
    ps.reset(input_state=st)
    out = ps.run(np.zeros(w + 15), output_form="dm")
    assert np.allclose(out, np.outer(st, st.conj()), atol=1e-10)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_stream_under_torchrun():
    n = 2 if torch.cuda.device_count() < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "tests", "multi_gpu_stream_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "STREAM_CHECK_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("fuse", [1, 3, 5])
def test_stream_long_pattern_hits_lane_slots_at_full_size(fuse):
    """Pattern several windows long: the low (lane-bit) slots are measured while the window is
    still full, which takes the warp-shuffle kernel (stream_lane_kernel)."""
    from mentpy_b200.streaming import LocalPass

    for name, args, w in (("linear_cluster", [60], 14), ("grid_cluster", [3, 12], 11)):
        gs = getattr(mb.templates, name)(*args)
        ang = np.random.default_rng(21).uniform(0, 2 * np.pi, len(gs.trainable_nodes))
        ps = mb.PatternSimulator(gs, backend="cuda-sv-stream", window_size=w, fuse=fuse)
        got = ps.run(ang, output_form="sv")
        want = matrix_free.run_sv_batch(PatternData.from_circuit(gs), ang, window_size=w)[0]
        assert infidelity_pure(got, want) < 1e-10
        assert np.allclose(got, want, atol=1e-9)
        assert any(isinstance(p, LocalPass) and p.lane for p in ps.simulator.last_schedule.passes)
