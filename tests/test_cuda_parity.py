"""GPU: parity of the CUDA path (through the C ABI) with the reference's golden vectors and with
the oracles on seeded inputs.  Tolerances: infidelity <= 1e-10 in fp64 (BASELINE.json
north_star); amplitude / matrix-element comparisons use 1e-9 where the reference's own global
phase is well conditioned."""
import numpy as np
import pytest
import torch

import mentpy_b200 as mb
from conftest import dm_distance, from_cplx, infidelity_pure, load_golden
from oracle import matrix_free
from oracle.pattern_data import PatternData

pytestmark = pytest.mark.gpu

INFID_TOL = 1e-10
CASES = load_golden("sim_cases.json")["cases"]


def _case_id(c):
    return f"{c['spec'][0]}{c['spec'][1]}-{c['backend']}-s{c['seed']}-w{c['window_size']}"


def _build(case):
    name, args, kwargs = case["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    for v in case["x_nodes"]:
        gs[v] = mb.Ment("X")
    for v, (ang, plane) in case["fixed"].items():
        gs[int(v)] = mb.Ment(ang, plane)
    return gs


def _dm_infidelity(rho, sigma):
    """1 - tr(rho sigma) for a pure reference sigma; max-abs distance otherwise is used."""
    return abs(1.0 - np.real(np.trace(rho @ sigma)))


@pytest.mark.parametrize("case", CASES, ids=[_case_id(c) for c in CASES])
def test_golden_cases_through_facade(case):
    gs = _build(case)
    inp = from_cplx(case["input_state"])
    backend = "cuda-sv" if case["backend"] == "numpy-sv" else "cuda-dm"
    ps = mb.PatternSimulator(gs, input_state=inp, backend=backend, window_size=case["window_size"])
    assert ps.window_size == case["window_size"]
    want = from_cplx(case["output"])
    ang = np.asarray(case["angles"])
    if backend == "cuda-sv":
        if "trace" in case:  # stateful measure() path, state after every measurement
            for node, ref_state in zip(ps.schedule_measure, case["trace"]):
                a = ang[gs.trainable_nodes.index(node)] if node in gs.trainable_nodes else gs[node].angle
                st, outcome = ps.measure(a)
                ref_state = from_cplx(ref_state)
                assert outcome == 0 and st.shape == ref_state.shape
                assert infidelity_pure(st, ref_state) < INFID_TOL
                assert np.allclose(st, ref_state, atol=1e-9, rtol=0)
            with pytest.raises(ValueError, match="No more measurements"):
                ps.measure(0.0)
            ps.reset()
        got = ps.run(ang, output_form=case["output_form"])
        assert got.shape == want.shape and got.dtype == np.complex128
        if case["output_form"] == "sv":
            assert infidelity_pure(got, want) < INFID_TOL
            assert np.allclose(got, want, atol=1e-9, rtol=0)  # including the reference's phase
        else:
            assert dm_distance(got, want) < 1e-10
        assert ps.outcomes == {v: 0 for v in ps.schedule_measure}
        with pytest.raises(ValueError, match="No more measurements"):
            ps.run(ang)
    else:
        got = ps.run(ang)
        assert got.shape == want.shape
        assert dm_distance(got, want) < 1e-10
        assert {str(k): v for k, v in ps.outcomes.items()} == case["outcomes"]


def test_default_output_is_density_matrix_and_call_alias():
    gs = mb.templates.grid_cluster(2, 4)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    ang = np.random.default_rng(3).uniform(0, 2 * np.pi, len(gs.trainable_nodes))
    rho = ps(ang)
    ps.reset()
    psi = ps.run(ang, output_form="statevector")
    assert rho.shape == (4, 4) and np.allclose(rho, np.outer(psi, psi.conj()), atol=1e-12)
    with pytest.raises(ValueError):
        ps.reset(); ps.run(ang[:-1])
    with pytest.raises(ValueError):
        ps.reset(); ps.run(ang, output_form="nope")


def test_teleportation_identity_all_backends():
    """Reference KAT tests/test_simulators.py:13-30 (atol 1e-3 there; 1e-10 here)."""
    from scipy.stats import unitary_group

    for backend in ("cuda-sv", "cuda-dm"):
        for i in range(1, 5):
            gs = mb.templates.linear_cluster(2 * i + 1)
            ps = mb.PatternSimulator(gs, backend=backend)
            for s in range(3):
                st = unitary_group.rvs(2, random_state=10 * i + s)[:, 0]
                ps.reset(input_state=st)
                assert len(ps.mbqcircuit.trainable_nodes) == 2 * i
                out = ps([0] * (2 * i))
                assert np.allclose(out, np.outer(st, st.conj()), atol=1e-10)


@pytest.mark.parametrize("spec,w", [(("linear_cluster", [5]), None), (("grid_cluster", [2, 6]), None),
                                    (("grid_cluster", [4, 5]), None), (("grid_cluster", [3, 5]), 5),
                                    (("grid_cluster", [2, 6]), 7), (("grid_cluster", [3, 6]), 9),
                                    (("linear_cluster", [16]), 12), (("many_wires", [[3, 4, 2]]), None),
                                    (("muta", [2, 1]), 5)])
def test_sv_batch_matches_oracle(spec, w):
    name, args = spec
    gs = getattr(mb.templates, name)(*args)
    pat = PatternData.from_circuit(gs)
    rng = np.random.default_rng(42)
    B, T = 257, len(gs.trainable_nodes)
    ang = rng.uniform(0, 2 * np.pi, (B, T))
    kw = {} if w is None else {"window_size": w}
    ps = mb.PatternSimulator(gs, backend="cuda-sv", **kw)
    got = ps.run_batch(ang)
    want = matrix_free.run_sv_batch(pat, ang, window_size=(w or 1))
    assert got.shape == want.shape
    infid = 1 - np.abs(np.sum(got.conj() * want, axis=1)) ** 2
    assert np.max(np.abs(infid)) < INFID_TOL
    # per-sample Haar inputs and the 'dm' form
    from scipy.stats import unitary_group

    n_in = len(gs.input_nodes)
    ins = np.stack([unitary_group.rvs(2**n_in, random_state=s)[:, 0] for s in range(8)])
    got = ps.run_batch(ang[:8], input_states=ins, output_form="dm")
    want = matrix_free.run_sv_batch(pat, ang[:8], ins, window_size=(w or 1), output_form="dm")
    assert dm_distance(got, want) < 1e-10
    # torch in -> torch out, no host round trip
    tout = ps.run_batch(torch.from_numpy(ang).cuda())
    assert isinstance(tout, torch.Tensor) and tout.is_cuda
    assert np.array_equal(tout.cpu().numpy(), ps.run_batch(ang))


@pytest.mark.parametrize("spec,w", [(("grid_cluster", [3, 8]), None), (("grid_cluster", [2, 5]), 5),
                                    (("linear_cluster", [6]), 3), (("grid_cluster", [2, 6]), 6),
                                    (("many_wires", [[3, 4, 2]]), None), (("linear_cluster", [4]), 1)])
def test_dm_batch_matches_oracle(spec, w):
    name, args = spec
    gs = getattr(mb.templates, name)(*args)
    pat = PatternData.from_circuit(gs)
    rng = np.random.default_rng(7)
    B, T = 37, len(gs.trainable_nodes)
    ang = rng.uniform(0, 2 * np.pi, (B, T))
    kw = {} if w is None else {"window_size": w}
    ps = mb.PatternSimulator(gs, backend="cuda-dm", **kw)
    got, oc = ps.run_batch(ang, return_outcomes=True)
    want, woc = matrix_free.run_dm_batch(pat, ang, window_size=(w or 1), return_outcomes=True)
    assert dm_distance(got, want) < 1e-10
    assert np.array_equal(oc, woc)
    assert np.allclose(np.trace(got, axis1=1, axis2=2), 1.0, atol=1e-12)


@pytest.mark.parametrize("kind,kw", [("depolarizing", {"p": 0.0}), ("depolarizing", {"p": 0.01}),
                                     ("depolarizing", {"p": 0.1}), ("amplitude_damping", {"p": 0.2}),
                                     ("phase_damping", {"p": 0.3}), ("phase_flip", {"p": 0.05}),
                                     ("generalized_amplitude_damping", {"p": 0.1, "p_gad": 0.3})])
def test_dm_noise_matches_oracles(kind, kw):
    """Noise parity is UNPINNED by the reference (PennyLane-only); compare with the windowed
    numpy restatement and the independent full-graph brute force."""
    from oracle import fullgraph_noise

    rng = np.random.default_rng(11)
    for name, args, w in (("grid_cluster", [3, 8], None), ("grid_cluster", [2, 4], 4), ("linear_cluster", [6], 3)):
        gs = getattr(mb.templates, name)(*args)
        pat = PatternData.from_circuit(gs)
        ang = rng.uniform(0, 2 * np.pi, (5, len(gs.trainable_nodes)))
        ps = mb.PatternSimulator(gs, backend="cuda-dm", circuit_noise=kind, **kw, **({} if w is None else {"window_size": w}))
        got = ps.run_batch(ang)
        okw = {"p": kw["p"]}
        if "p_gad" in kw:
            okw["p_gad"] = kw["p_gad"]
        want = matrix_free.run_dm_batch(pat, ang, window_size=(w or 1), noise=kind, noise_kwargs=okw)
        assert dm_distance(got, want) < 1e-10
        assert np.allclose(np.trace(got, axis1=1, axis2=2), 1.0, atol=1e-12)
        assert np.allclose(got, np.conj(np.swapaxes(got, 1, 2)), atol=1e-12)
        if pat.n_nodes <= 8:
            brute = fullgraph_noise.run_fullgraph_dm(pat, ang[0], noise=kind, noise_kwargs=okw)
            assert dm_distance(got[0], brute) < 1e-10
    if kw.get("p") == 0.0:
        clean = mb.PatternSimulator(gs, backend="cuda-dm", window_size=w).run_batch(ang)
        assert dm_distance(got, clean) < 1e-13


def test_depolarizing_three_quarters_is_maximally_mixed():
    gs = mb.templates.linear_cluster(4)
    ps = mb.PatternSimulator(gs, backend="cuda-dm", circuit_noise="depolarizing", p=0.75)
    rho = ps.run_batch(np.random.default_rng(0).uniform(0, 6, (3, 3)))
    assert np.allclose(rho, np.eye(2) / 2, atol=1e-12)


def test_dm_outcome1_quirk():
    d = load_golden("dm_outcome_quirk.json")
    name, args, kwargs = d["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    ps = mb.PatternSimulator(gs, input_state=from_cplx(d["input_state"]), backend="cuda-dm",
                             window_size=d["window_size"])
    for run in d["runs"]:
        ps.reset()
        got = ps.run(np.asarray(run["angles"]))
        assert dm_distance(got, from_cplx(run["output"])) < 1e-10
        assert {str(k): v for k, v in ps.outcomes.items()} == run["outcomes"]
        assert ps.last_status[0] & 2 == (2 if 1 in run["outcomes"].values() else 0)


def test_dm_stateful_measure_and_planes():
    case = next(c for c in CASES if c["backend"] == "numpy-dm" and c["fixed"] and "XZ" in str(c["fixed"]))
    gs = _build(case)
    ps = mb.PatternSimulator(gs, backend="cuda-dm", window_size=case["window_size"])
    ang = np.asarray(case["angles"])
    for node in ps.schedule_measure:
        a = ang[gs.trainable_nodes.index(node)] if node in gs.trainable_nodes else gs[node].angle
        rho, outcome = ps.measure(a)
        assert abs(np.trace(rho) - 1) < 1e-12
    k = len(gs.output_nodes)
    assert rho.shape == (2**k, 2**k)
    # after the last measurement the window is the output block in schedule order
    final = ps.reorder_qubits(rho, ps.current_simulated_nodes(), gs.quantum_output_nodes)
    assert dm_distance(final, from_cplx(case["output"])) < 1e-10
    with pytest.raises(ValueError, match="fixed angle"):
        ps.reset()
        fixed_first = next(i for i, n in enumerate(ps.schedule_measure) if n not in gs.trainable_nodes)
        for n in ps.schedule_measure[:fixed_first]:
            ps.measure(0.1)
        ps.measure(123.0)


def test_full_size_properties_c2():
    """BASELINE configs[1] at full size: 65,536 angle sets on grid_cluster(2,6); size-independent
    checks: norms, batch-order independence, agreement of a strided subsample with the oracle."""
    gs = mb.templates.grid_cluster(2, 6)
    pat = PatternData.from_circuit(gs)
    B = 65536
    ang = np.random.default_rng(1).uniform(0, 2 * np.pi, (B, 10))
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    out = ps.run_batch(ang)
    assert out.shape == (B, 4)
    assert np.allclose(np.sum(np.abs(out) ** 2, axis=1), 1.0, atol=1e-12)
    perm = np.random.default_rng(2).permutation(B)
    assert np.array_equal(ps.run_batch(ang[perm]), out[perm])
    idx = np.arange(0, B, 97)
    want = matrix_free.run_sv_batch(pat, ang[idx])
    infid = 1 - np.abs(np.sum(out[idx].conj() * want, axis=1)) ** 2
    assert np.max(np.abs(infid)) < INFID_TOL
    # golden seed-1 vector is row 0 of this very batch
    case = next(c for c in CASES if c["spec"][1] == [2, 6] and c["output_form"] == "sv")
    assert np.allclose(out[0], from_cplx(case["output"]), atol=1e-9)


def test_gradient_kernel_matches_reference_golden():
    from mentpy_b200.gradients import psr_gradient_batched

    g = load_golden("gradients.json")["c4"]
    gs = mb.templates.grid_cluster(4, 5)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    x = np.asarray(g["x"])
    grad, cost = psr_gradient_batched(ps, x[None, :], from_cplx(g["target"]), return_cost=True)
    assert abs(cost[0] - g["cost"]) < 1e-12
    assert np.allclose(grad[0], g["psr"], atol=1e-11, rtol=0)
    grad_fd = psr_gradient_batched(ps, x[None, :], from_cplx(g["target"]), shift=1e-5)
    assert np.allclose(grad_fd[0], g["fd"], atol=1e-6, rtol=0)


def test_host_pipeline_pinned_and_pageable_agree():
    """run_batch on host arrays goes through mbqc_run_batch_sv_host (chunked, multi-stream)."""
    gs = mb.templates.grid_cluster(2, 6)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    B = 20000  # several chunks, ragged last chunk
    ang = np.random.default_rng(8).uniform(0, 2 * np.pi, (B, 10))
    dev_out = ps.run_batch(torch.from_numpy(ang).cuda()).cpu().numpy()
    pinned = mb.pinned_empty((B, 10))
    pinned.copy_(torch.from_numpy(ang))
    a = ps.run_batch(ang)
    b = ps.run_batch(pinned, copy=False)
    assert np.array_equal(a, dev_out) and np.array_equal(b, dev_out)
    c = ps.run_batch(pinned, output_form="dm")
    assert c.shape == (B, 4, 4)
    assert np.allclose(c, dev_out[:, :, None] * dev_out.conj()[:, None, :], atol=1e-14)
    # tiny and empty batches
    assert ps.run_batch(ang[:1]).shape == (1, 4)
    assert ps.run_batch(np.zeros((0, 10))).shape == (0, 4)


def test_long_pattern_unstaged_angles_and_renormalisation():
    """T large enough that the (cos,sin) tile does not fit shared memory -> global-angle fallback;
    also exercises the periodic renormalisation (> 16 steps)."""
    gs = mb.templates.linear_cluster(60)
    pat = PatternData.from_circuit(gs)
    ang = np.random.default_rng(12).uniform(0, 2 * np.pi, (64, 59))
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    got = ps.run_batch(ang)
    want = matrix_free.linear_cluster_analytic(ang)
    infid = 1 - np.abs(np.sum(got.conj() * want, axis=1)) ** 2
    assert np.max(np.abs(infid)) < INFID_TOL
    gs2 = mb.templates.grid_cluster(2, 40)
    ang2 = np.random.default_rng(13).uniform(0, 2 * np.pi, (16, 78))
    got2 = mb.PatternSimulator(gs2, backend="cuda-sv").run_batch(ang2)
    want2 = matrix_free.run_sv_batch(PatternData.from_circuit(gs2), ang2)
    infid = 1 - np.abs(np.sum(got2.conj() * want2, axis=1)) ** 2
    assert np.max(np.abs(infid)) < INFID_TOL


def test_batched_gradient_random_patterns():
    from mentpy_b200.gradients import psr_gradient_batched
    from scipy.stats import unitary_group

    for name, args in (("grid_cluster", [2, 4]), ("linear_cluster", [6]), ("grid_cluster", [3, 4])):
        gs = getattr(mb.templates, name)(*args)
        pat = PatternData.from_circuit(gs)
        T, k = len(gs.trainable_nodes), len(gs.output_nodes)
        X = np.random.default_rng(3).uniform(0, 2 * np.pi, (9, T))
        tgt = unitary_group.rvs(2**k, random_state=4)[:, 0]
        ps = mb.PatternSimulator(gs, backend="cuda-sv")
        grad, cost = psr_gradient_batched(ps, X, tgt, return_cost=True)

        def cost_fn(x):
            psi = matrix_free.run_sv_batch(pat, x)[0]
            return 1 - abs(np.vdot(tgt, psi)) ** 2

        for b in range(3):
            assert abs(cost[b] - cost_fn(X[b])) < 1e-12
            for i in range(T):
                e = np.zeros(T); e[i] = 1.5
                want = (cost_fn(X[b] + e) - cost_fn(X[b] - e)) / 3.0
                assert abs(grad[b, i] - want) < 1e-11


@pytest.mark.parametrize("backend", ["cuda-sv", "cuda-dm"])
def test_slot_layout_independence(backend):
    """Results must not depend on which bit slot a qubit lives in: 'lsb' slot order drives the
    register kernel through its generic (non-unrolled) slot path."""
    rng = np.random.default_rng(21)
    for name, args, w in (("grid_cluster", [2, 6], None), ("grid_cluster", [3, 5], 5), ("linear_cluster", [7], 3)):
        gs = getattr(mb.templates, name)(*args)
        gs[1] = mb.Ment("X")
        gs[2] = mb.Ment(0.77, "XY")
        ang = rng.uniform(0, 2 * np.pi, (33, len(gs.trainable_nodes)))
        kw = {} if w is None else {"window_size": w}
        a = mb.PatternSimulator(gs, backend=backend, **kw).run_batch(ang)
        b = mb.PatternSimulator(gs, backend=backend, slot_order="lsb", **kw).run_batch(ang)
        assert np.allclose(a, b, atol=1e-13, rtol=0)
        pat = PatternData.from_circuit(gs)
        want = (matrix_free.run_sv_batch if backend == "cuda-sv" else matrix_free.run_dm_batch)(pat, ang, window_size=(w or 1))
        assert np.allclose(a, want, atol=1e-10, rtol=0)


def test_sincos_accuracy_over_wide_angle_range():
    """The kernel's own sincos (Cody-Waite + fdlibm kernels) against numpy over many periods,
    through linear_cluster(2): output ~ J(-theta)|+> = (1 + e^{-i theta}, 1 - e^{-i theta})/2."""
    gs = mb.templates.linear_cluster(3)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    th = np.concatenate([np.linspace(-50, 50, 4001), np.array([1e4 + 0.3, -9.9e4, 2e5, 1e7 + 0.1, 0.0, np.pi, -np.pi / 2])])
    ang = np.stack([th, np.zeros_like(th)], axis=1)
    got = ps.run_batch(ang)
    want = matrix_free.linear_cluster_analytic(ang)
    infid = 1 - np.abs(np.sum(got.conj() * want, axis=1)) ** 2
    assert np.max(np.abs(infid)) < 1e-13


def test_full_size_properties_c3_c4():
    """BASELINE configs 3 and 4 at full per-GPU size through size-independent properties."""
    from mentpy_b200.gradients import psr_gradient_batched

    # C3: 4,096 angle sets, grid_cluster(3,8) DM with depolarizing noise
    gs = mb.templates.grid_cluster(3, 8)
    ang = np.random.default_rng(2).uniform(0, 2 * np.pi, (4096, 21))
    clean = mb.PatternSimulator(gs, backend="cuda-dm").run_batch(ang)
    noisy = mb.PatternSimulator(gs, backend="cuda-dm", circuit_noise="depolarizing", p=0.01).run_batch(ang)
    for rho in (clean, noisy):
        assert rho.shape == (4096, 8, 8)
        assert np.allclose(np.trace(rho, axis1=1, axis2=2), 1.0, atol=1e-12)
        assert np.allclose(rho, np.conj(np.swapaxes(rho, 1, 2)), atol=1e-12)
    purity = np.real(np.einsum("bij,bji->b", clean, clean))
    assert np.allclose(purity, 1.0, atol=1e-10)                      # noiseless run stays pure
    assert np.all(np.real(np.einsum("bij,bji->b", noisy, noisy)) < 1.0 - 1e-3)
    case = next(c for c in CASES if c["spec"][1] == [3, 8])          # golden seed-2 row
    assert dm_distance(clean[0], from_cplx(case["output"])) < 1e-10
    # SV and DM backends agree on the whole batch (reference: tests/test_simulators.py:33-64)
    sv = mb.PatternSimulator(gs, backend="cuda-sv").run_batch(ang[:512], output_form="dm")
    assert dm_distance(sv, clean[:512]) < 1e-10

    # C4: 65,536 base vectors x 32 shifted evaluations, grid_cluster(4,5)
    gs = mb.templates.grid_cluster(4, 5)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    X = np.random.default_rng(4).uniform(0, 2 * np.pi, (65536, 16))
    tgt = np.full(16, 0.25)
    g, c = psr_gradient_batched(ps, X, tgt, return_cost=True)
    assert g.shape == (65536, 16) and np.all(np.isfinite(g)) and np.all((c > -1e-12) & (c < 1 + 1e-12))
    g2 = psr_gradient_batched(ps, X + 2 * np.pi, tgt)                # 2 pi periodicity in every angle
    assert np.allclose(g, g2, atol=1e-9)
    gold = load_golden("gradients.json")["c4"]                       # row 0 is the golden vector
    assert np.allclose(g[0], gold["psr"], atol=1e-11) and abs(c[0] - gold["cost"]) < 1e-12
    # explicit shifted evaluations on a subsample
    idx = np.arange(0, 65536, 4099)
    for i in (0, 7, 15):
        e = np.zeros(16); e[i] = 1.5
        fp = 1 - np.abs(ps.run_batch(X[idx] + e) @ tgt) ** 2
        fm = 1 - np.abs(ps.run_batch(X[idx] - e) @ tgt) ** 2
        assert np.allclose(g[idx, i], (fp - fm) / 3.0, atol=1e-11)


def test_edge_cases_sizes_and_shapes():
    """Smallest / largest windows of every batched kernel, patterns without trainable angles,
    batch sizes around CTA boundaries."""
    rng = np.random.default_rng(31)
    # window 1..2 (register kernel W=1,2), including the all-fixed-angle pattern (T = 0)
    gs = mb.templates.linear_cluster(4)
    for v in (0, 1, 2):
        gs[v] = mb.Ment("X") if v != 1 else mb.Ment(0.9, "XY")
    assert gs.trainable_nodes == []
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    got = ps.run_batch(np.zeros((5, 0)))
    want = matrix_free.run_sv_batch(PatternData.from_circuit(gs), np.zeros((5, 0)))
    assert np.allclose(got, want, atol=1e-12)
    assert np.allclose(ps.run([]), np.outer(want[0], want[0].conj()), atol=1e-12)
    pd = mb.PatternSimulator(gs, backend="cuda-dm")
    assert np.allclose(pd.run([]), np.outer(want[0], want[0].conj()), atol=1e-12)
    # batch sizes straddling the 128-thread CTA
    gs = mb.templates.grid_cluster(2, 5)
    pat = PatternData.from_circuit(gs)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    for B in (1, 127, 128, 129, 255, 257):
        ang = rng.uniform(0, 2 * np.pi, (B, 8))
        assert np.allclose(ps.run_batch(torch.from_numpy(ang).cuda()).cpu().numpy(),
                           matrix_free.run_sv_batch(pat, ang), atol=1e-10)
    # strided (non-contiguous rows) device input
    big = torch.from_numpy(rng.uniform(0, 2 * np.pi, (64, 16))).cuda()
    view = big[:, :8]
    assert np.allclose(ps.run_batch(view).cpu().numpy(), matrix_free.run_sv_batch(pat, view.cpu().numpy()), atol=1e-10)
    # largest shared-memory windows: SV w = 12, DM w = 6
    gs = mb.templates.linear_cluster(18)
    ang = rng.uniform(0, 2 * np.pi, (3, 17))
    got = mb.PatternSimulator(gs, backend="cuda-sv", window_size=12).run_batch(ang)
    assert np.max(np.abs(1 - np.abs(np.sum(got.conj() * matrix_free.linear_cluster_analytic(ang), axis=1)) ** 2)) < 1e-10
    gs = mb.templates.grid_cluster(2, 6)
    ang = rng.uniform(0, 2 * np.pi, (3, 10))
    got = mb.PatternSimulator(gs, backend="cuda-dm", window_size=6).run_batch(ang)
    assert dm_distance(got, matrix_free.run_dm_batch(PatternData.from_circuit(gs), ang, window_size=6)) < 1e-10
    with pytest.raises(NotImplementedError):
        mb.PatternSimulator(gs, backend="cuda-dm", window_size=7).run_batch(ang)
    with pytest.raises(NotImplementedError):
        mb.PatternSimulator(mb.templates.linear_cluster(20), backend="cuda-sv", window_size=13).run_batch(np.zeros((1, 19)))
    # wrong shapes
    with pytest.raises(ValueError):
        ps.run_batch(np.zeros((4, 7)))
    with pytest.raises(ValueError):
        ps.run_batch(np.zeros((4, 8)), input_states=np.zeros((3, 4)))
    with pytest.raises(ValueError):
        mb.PatternSimulator(mb.templates.grid_cluster(2, 5), input_state=np.ones(3), backend="cuda-sv").run(np.zeros(8))


@pytest.mark.parametrize("kernel", ["pairs", "prefix"])
def test_gradient_kernels_agree(kernel, monkeypatch):
    """Both gradient kernels (thread per (vector, parameter) / thread per vector with prefix
    sharing) against explicit shifted evaluations, incl. fixed-angle and X nodes, Haar inputs."""
    from mentpy_b200.gradients import psr_gradient_batched
    from scipy.stats import unitary_group

    monkeypatch.setenv("MBQC_GRAD_KERNEL", kernel)
    for name, args, w in (("grid_cluster", [4, 5], None), ("grid_cluster", [2, 7], None), ("linear_cluster", [40], 4)):
        gs = getattr(mb.templates, name)(*args)
        gs[1] = mb.Ment("X")
        gs[2] = mb.Ment(0.37, "XY")
        T, k, n_in = len(gs.trainable_nodes), len(gs.output_nodes), len(gs.input_nodes)
        ps = mb.PatternSimulator(gs, backend="cuda-sv", **({} if w is None else {"window_size": w}))
        X = np.random.default_rng(6).uniform(0, 2 * np.pi, (37, T))
        tgt = unitary_group.rvs(2**k, random_state=8)[:, 0]
        ins = np.stack([unitary_group.rvs(2**n_in, random_state=s)[:, 0] for s in range(37)])
        for inputs in (None, ins):
            g, c = psr_gradient_batched(ps, X, tgt, input_states=inputs, return_cost=True)
            f = lambda Y: 1 - np.abs(ps.run_batch(Y, input_states=inputs) @ tgt.conj()) ** 2
            assert np.allclose(c, f(X), atol=1e-12)
            for i in (0, T // 2, T - 1):
                e = np.zeros(T); e[i] = 1.5
                assert np.allclose(g[:, i], (f(X + e) - f(X - e)) / 3.0, atol=1e-11)


def test_complex64_mode():
    """Optional complex64 mode: infidelity <= 1e-5 vs the fp64 reference (BASELINE north_star)."""
    rng = np.random.default_rng(41)
    for name, args, w in (("linear_cluster", [5], None), ("grid_cluster", [2, 6], None),
                          ("grid_cluster", [4, 5], None), ("grid_cluster", [3, 5], 5), ("linear_cluster", [40], 3)):
        gs = getattr(mb.templates, name)(*args)
        gs[1] = mb.Ment("X")
        pat = PatternData.from_circuit(gs)
        T = len(gs.trainable_nodes)
        ang = rng.uniform(0, 2 * np.pi, (300, T))
        kw = {} if w is None else {"window_size": w}
        ps = mb.PatternSimulator(gs, backend="cuda-sv", dtype="complex64", **kw)
        got = ps.run_batch(ang)
        assert got.dtype == np.complex64
        want = matrix_free.run_sv_batch(pat, ang, window_size=(w or 1))
        infid = np.abs(1 - np.abs(np.sum(got.astype(np.complex128).conj() * want, axis=1)) ** 2)
        assert infid.max() < 1e-5, infid.max()
        assert np.allclose(np.sum(np.abs(got.astype(np.complex128)) ** 2, axis=1), 1.0, atol=1e-5)
        rho = ps.run_batch(ang[:7], output_form="dm")
        assert rho.dtype == np.complex64 and np.abs(rho - want[:7, :, None] * want[:7, None, :].conj()).max() < 1e-4
        tout = ps.run_batch(torch.from_numpy(ang).cuda())
        assert tout.dtype == torch.complex64 and tout.is_cuda
    # golden C2 vector through the single-sample API
    case = next(c for c in CASES if c["spec"][1] == [2, 6] and c["output_form"] == "sv")
    ps = mb.PatternSimulator(_build(case), backend="cuda-sv", dtype="complex64")
    out = ps.run(np.asarray(case["angles"]), output_form="sv")
    assert infidelity_pure(out.astype(np.complex128), from_cplx(case["output"])) < 1e-5
    with pytest.raises(NotImplementedError):
        mb.PatternSimulator(mb.templates.linear_cluster(12), backend="cuda-sv", dtype="complex64",
                            window_size=8).run_batch(np.zeros((1, 11)))
    with pytest.raises(ValueError):
        mb.PatternSimulator(gs, backend="cuda-sv", dtype="float16")


def test_calculator_helpers_match_reference():
    """mentpy tests/test_calculator.py:12-54 restated + golden vectors from the reference
    (tests/golden/helpers.json) for the CUDA calculator kernels."""
    h = load_golden("helpers.json")
    calc = mb.calculator
    psi = from_cplx(h["sum_trace_pure"]["psi"])
    assert np.allclose(calc.partial_trace(psi, [0]), from_cplx(h["sum_trace_pure"]["idx0"]), atol=1e-13)
    assert np.allclose(calc.partial_trace(psi, [1]), from_cplx(h["sum_trace_pure"]["idx1"]), atol=1e-13)
    rho = calc.pure2density(psi)
    assert np.allclose(rho, np.outer(psi, psi.conj()), atol=1e-15)
    assert np.allclose(calc.partial_trace(rho, [0]), from_cplx(h["trace_mixed"]["idx0"]), atol=1e-13)
    assert np.allclose(calc.partial_trace(rho, [2]), from_cplx(h["trace_mixed"]["idx2"]), atol=1e-13)
    # reference unit tests
    assert np.allclose(calc.pure2density(np.array([1, 0])), [[1, 0], [0, 0]])
    plus = np.array([1, 1]) / np.sqrt(2)
    assert np.allclose(calc.pure2density(plus), [[0.5, 0.5], [0.5, 0.5]])
    mixed = np.array([[0.5, 0], [0, 0.5]])
    sigma = np.outer(plus, plus)
    prod = np.kron(mixed, sigma)
    assert np.allclose(calc.partial_trace(mixed, [0]), 1)
    assert np.allclose(calc.partial_trace(prod, [0]), sigma) and np.allclose(calc.partial_trace(prod, [1]), mixed)
    assert np.allclose(calc.partial_trace(plus, [0]), 1)
    pp = np.kron(plus, np.array([1.0, 0.0]))
    assert np.allclose(calc.partial_trace(pp, [0]), [1, 0]) and np.allclose(calc.partial_trace(pp, [1]), plus)
    # multi-qubit traces on a 5-qubit state against numpy
    st = mb.utils.generate_haar_random_states(5, 1, random_state=3)[0]
    t = st.reshape([2] * 5).transpose(0, 2, 4, 1, 3).reshape(8, 4).sum(axis=1)
    assert np.allclose(calc.partial_trace(st, [1, 3]), t / np.linalg.norm(t), atol=1e-13)
    r5 = np.outer(st, st.conj()).reshape([2] * 10)
    want = np.einsum("abcdeAbCdE->aceACE", r5).reshape(8, 8)
    assert np.allclose(calc.partial_trace(np.outer(st, st.conj()), [1, 3]), want, atol=1e-13)
    with pytest.raises(ValueError):
        calc.partial_trace(np.zeros((2, 2, 2)), [0])
    with pytest.raises(ValueError):
        calc.partial_trace(psi, [7])


def test_dm_kernels_agree(monkeypatch):
    """The register/shuffle DM kernel (w <= 5) and the shared-memory DM kernel give the same
    states and outcome records (incl. noise, XZ/YZ planes, Haar inputs, outcome-1 branch)."""
    from scipy.stats import unitary_group

    rng = np.random.default_rng(51)
    for name, args, w in (("grid_cluster", [3, 8], None), ("grid_cluster", [2, 5], 4), ("grid_cluster", [4, 4], None), ("linear_cluster", [6], 3),
                          ("linear_cluster", [5], None), ("many_wires", [[3, 3]], 2)):
        gs = getattr(mb.templates, name)(*args)
        if name == "grid_cluster":
            gs[1] = mb.Ment(0.7, "XZ")
            gs[2] = mb.Ment("YZ")
        T, n_in = len(gs.trainable_nodes), len(gs.input_nodes)
        ang = rng.uniform(0, 2 * np.pi, (19, T))
        ang[0, 0] = np.pi                       # pushes the first measurement towards prob0 = 0 for |+> inputs
        ins = np.stack([unitary_group.rvs(2**n_in, random_state=s)[:, 0] for s in range(19)])
        kw = {} if w is None else {"window_size": w}
        for noise in ({}, {"circuit_noise": "amplitude_damping", "p": 0.15}):
            res = {}
            for kern in ("reg", "smem"):
                monkeypatch.setenv("MBQC_DM_KERNEL", kern)
                ps = mb.PatternSimulator(gs, backend="cuda-dm", **kw, **noise)
                res[kern] = ps.run_batch(ang, input_states=ins, return_outcomes=True)
                res[kern + "_plus"] = ps.run_batch(ang, return_outcomes=True)
            for suffix in ("", "_plus"):
                assert np.allclose(res["reg" + suffix][0], res["smem" + suffix][0], atol=1e-12)
                assert np.array_equal(res["reg" + suffix][1], res["smem" + suffix][1])


def test_async_host_calls_match_blocking_calls():
    """run_batch_async (submit / wait tickets): several calls in flight return exactly what the
    blocking call returns; too many outstanding calls and double waits are reported."""
    gs = mb.templates.grid_cluster(2, 6)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    rng = np.random.default_rng(12)
    batches = [rng.uniform(0, 2 * np.pi, (20000, 10)) for _ in range(7)]
    want = [ps.run_batch(a) for a in batches]
    pend, got = [], []
    for a in batches:
        pend.append(ps.run_batch_async(a))
        if len(pend) == 2:
            got.append(pend.pop(0).result(copy=True))
    got += [h.result(copy=True) for h in pend]
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    h = ps.run_batch_async(batches[0], output_form="dm")
    assert np.allclose(h.result(), want[0][:, :, None] * want[0].conj()[:, None, :], atol=1e-15)
    assert h.result() is not None and h.status_any == 0          # second result(): no second wait
    assert ps.run_batch_async(np.zeros((0, 10))).result().shape == (0, 4)
    hs = [ps.run_batch_async(batches[i]) for i in range(3)]       # rotating buffers: 3 sets
    with pytest.raises(ValueError):
        import torch
        ps.run_batch_async(torch.zeros((4, 10), device="cuda", dtype=torch.float64))
    for x in hs:
        x.result()
    lib = mb._lib.load()
    import ctypes as C
    assert lib.mbqc_host_wait(0, None) == mb._lib.MBQC_E_ARG and b"not in flight" in lib.mbqc_last_error()


def test_plane_z_expectation_mode():
    """Plane-Z nodes in mode='expectation' (np_simulator_dm.py:327-344): the qubit is traced out
    and prob1 recorded -- against the reference (tests/golden/dm_z_expectation.json), the oracle
    (batch, noise) and the stateful API."""
    from scipy.stats import unitary_group

    for c in load_golden("dm_z_expectation.json")["cases"]:
        name, args, kwargs = c["spec"]
        gs = getattr(mb.templates, name)(*args, **kwargs)
        for v, pl in c["planes"].items():
            gs[int(v)] = mb.Ment(pl)
        inp = None if c["input_state"] is None else from_cplx(c["input_state"])
        ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=c["window_size"])
        ang = np.asarray(c["angles"])
        rho = ps.run(ang, mode="expectation")
        want = from_cplx(c["output"])
        assert rho.shape == want.shape and np.abs(rho - want).max() < 1e-12
        assert {str(k): v for k, v in ps.outcomes.items()}.keys() == c["outcomes"].keys()
        for k, v in ps.outcomes.items():
            assert abs(v - c["outcomes"][str(k)]) < 1e-12
            assert isinstance(v, float) == (c["planes"].get(str(k)) == "Z")
        ps.reset()
        for node in ps.schedule_measure:                  # step-by-step API
            a = ang[gs.trainable_nodes.index(node)] if node in gs.trainable_nodes else None
            st, oc = ps.measure(a, mode="expectation")
            assert abs(oc - c["outcomes"][str(node)]) < 1e-12
        assert np.abs(st - want).max() < 1e-12
        # batch + noise against the oracle
        pat = PatternData.from_circuit(gs)
        B, n_in = 33, len(gs.input_nodes)
        A = np.random.default_rng(3).uniform(0, 2 * np.pi, (B, len(ang)))
        ins = np.stack([unitary_group.rvs(2**n_in, random_state=s)[:, 0] for s in range(B)])
        for noise in ({}, {"circuit_noise": "amplitude_damping", "p": 0.2}, {"circuit_noise": "depolarizing", "p": 0.1}):
            pn = mb.PatternSimulator(gs, backend="cuda-dm", window_size=c["window_size"], **noise)
            got, oc = pn.run_batch(A, input_states=ins, return_outcomes=True, mode="expectation")
            kw = {k: v for k, v in noise.items() if k != "circuit_noise"}
            ref, roc = matrix_free.run_dm_batch(pat, A, input_states=ins, window_size=c["window_size"],
                                                noise=noise.get("circuit_noise"), noise_kwargs=kw,
                                                return_outcomes=True, mode="expectation")
            assert np.abs(got - ref).max() < 1e-11 and np.abs(oc - roc).max() < 1e-11
    with pytest.raises(ValueError):
        mb.PatternSimulator(gs, backend="cuda-sv")        # SV path: XY planes only, like the reference


def test_xyz_plane_matches_reference_golden():
    """Fixed two-angle XYZ-plane nodes (ment.py:239-251) on every density-matrix kernel -- specialised
    (dm_jit_src.inc), register (dm_reg.cuh), shared-memory (dm_batch.cuh, window 6) -- against
    outputs recorded from the reference; with a channel against the oracle.  A trainable XYZ node
    raises TypeError like the reference (one float arrives where a tuple is needed)."""
    import warnings

    from mentpy_b200 import _lib

    lib = _lib.load()
    for c in load_golden("dm_xyz_plane.json")["cases"]:
        name, args, kwargs = c["spec"]
        gs = getattr(mb.templates, name)(*args, **kwargs)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for v, ang2 in c["xyz"].items():
                gs[int(v)] = mb.Ment(tuple(ang2), "XYZ")
        inp = None if c["input_state"] is None else from_cplx(c["input_state"])
        ang = np.asarray(c["angles"])
        want = from_cplx(c["output"])
        pat = PatternData.from_circuit(gs)
        rows = np.random.default_rng(c["seed"]).uniform(0, 2 * np.pi, (19, len(ang)))
        rows[7] = ang
        for mode in (0, 2):
            prev = lib.mbqc_jit_set_mode(mode)
            try:
                ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=c["window_size"])
                got = ps.run(ang)
                assert got.shape == want.shape and dm_distance(got, want) < 1e-10
                assert {str(k): v for k, v in ps.outcomes.items()} == c["outcomes"]
                batch = ps.run_batch(rows)
                assert dm_distance(batch[7], want) < 1e-10
                pn = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=c["window_size"],
                                         circuit_noise="depolarizing", p=0.05)
                ref = matrix_free.run_dm_batch(pat, rows[:5], input_states=None if inp is None else np.tile(inp, (5, 1)),
                                               window_size=c["window_size"], noise="depolarizing", noise_kwargs={"p": 0.05})
                assert dm_distance(pn.run_batch(rows[:5]), ref) < 1e-10
            finally:
                lib.mbqc_jit_set_mode(prev)
    assert "failures=0" in lib.mbqc_jit_info().decode()
    gs = mb.templates.grid_cluster(2, 4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gs[2] = mb.Ment("XYZ")
    with pytest.raises(TypeError, match="Expected tuple"):
        mb.PatternSimulator(gs, backend="cuda-dm")


def test_plane_z_sample_mode():
    """Plane-Z nodes in mode='sample' (np_simulator_dm.py:329-346): the reference draws their outcome
    from (prob0, prob1) even under force0 and projects.  Parity is statistical by nature (np.random
    there, Philox here): every shot's state must be the oracle's branch state for the outcomes the
    kernel recorded, the outcome frequencies must follow the branch probabilities, and a
    (seed, sample_offset) pair must reproduce the run."""
    for c in load_golden("dm_z_expectation.json")["cases"][:3]:
        name, args, kwargs = c["spec"]
        gs = getattr(mb.templates, name)(*args, **kwargs)
        for v, pl in c["planes"].items():
            gs[int(v)] = mb.Ment(pl)
        inp = None if c["input_state"] is None else from_cplx(c["input_state"])
        pat = PatternData.from_circuit(gs)
        ang = np.asarray(c["angles"])
        for noise in ({}, {"circuit_noise": "depolarizing", "p": 0.1}):
            ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=c["window_size"], seed=5, **noise)
            B = 4000
            rows = np.tile(ang, (B, 1))
            rho, oc = ps.run_batch(rows, return_outcomes=True, mode="sample", seed=5, sample_offset=0)
            assert oc.dtype == np.int8 and set(np.unique(oc)) <= {0, 1}
            kw = {k: v for k, v in noise.items() if k != "circuit_noise"}
            ins = None if inp is None else np.tile(inp, (64, 1))
            ref, roc, zp1 = matrix_free.run_dm_batch(pat, rows[:64], input_states=ins, window_size=c["window_size"],
                                                     noise=noise.get("circuit_noise"), noise_kwargs=kw, return_outcomes=True,
                                                     mode="sample", z_outcomes=oc[:64])
            assert dm_distance(rho[:64], ref) < 1e-10
            assert np.array_equal(roc.astype(np.int8), oc[:64])
            # first plane-Z step: every shot sees the same prob1 -> binomial frequency within 5 sigma
            zcols = [m for m, st in enumerate(ps.simulator.plan.steps) if st.plane == mb._lib.PLANE_Z]
            p1 = zp1[0, zcols[0]]
            freq = oc[:, zcols[0]].mean()
            assert abs(freq - p1) < 5 * np.sqrt(max(p1 * (1 - p1), 1e-4) / B) + 1e-3
            again, oc2 = ps.run_batch(rows[:100], return_outcomes=True, mode="sample", seed=5, sample_offset=0)
            assert np.array_equal(oc2, oc[:100]) and np.array_equal(again, rho[:100])
            other, oc3 = ps.run_batch(rows[:100], return_outcomes=True, mode="sample", seed=5, sample_offset=100)
            assert np.array_equal(oc3, oc[100:200])
        # stateful API: one shot, prefix re-runs repeat the earlier draws
        ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=c["window_size"], seed=11)
        ps.reset()
        rec = []
        for node in ps.schedule_measure:
            a = ang[gs.trainable_nodes.index(node)] if node in gs.trainable_nodes else None
            st, o = ps.measure(a)
            rec.append(int(o))
        ins1 = None if inp is None else inp[None]
        ref1, _, _ = matrix_free.run_dm_batch(pat, ang[None], input_states=ins1, window_size=c["window_size"],
                                              return_outcomes=True, mode="sample", z_outcomes=np.asarray(rec)[None])
        assert dm_distance(st, ref1[0]) < 1e-10


def test_controlled_measurements_match_reference_golden():
    """Outcome-controlled measurements (operators/controlled_ment.py:14-113) on the density-matrix
    kernels (register kernel for window <= 5, shared-memory kernel at window 6): single runs against
    outputs recorded from the reference -- including the condition fired by a real outcome 1 --,
    batches and a channel against the oracle, and the stateful API."""
    import warnings

    from oracle.gen_golden import CONTROL_CASES

    for c in load_golden("dm_controlled.json")["cases"]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gs = CONTROL_CASES[c["name"]](mb, mb.ControlMent)
        inp = None if c["input_state"] is None else from_cplx(c["input_state"])
        ang = np.asarray(c["angles"])
        want = from_cplx(c["output"])
        pat = PatternData.from_circuit(gs)
        windows = [c["window_size"]] + ([6] if len(gs.graph.nodes()) >= 8 and c["name"] != "real_outcome_1" else [])
        for w in windows:
            ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=w)
            got = ps.run(ang)
            if w == c["window_size"]:
                assert got.shape == want.shape and dm_distance(got, want) < 1e-10, c["name"]
                assert {str(k): v for k, v in ps.outcomes.items()} == c["outcomes"]
            rows = np.random.default_rng(c["seed"]).uniform(0, 2 * np.pi, (23, len(ang)))
            rows[3] = ang
            ins = None if inp is None else np.tile(inp, (23, 1))
            for noise in ({}, {"circuit_noise": "amplitude_damping", "p": 0.15}):
                pn = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=w, **noise)
                batch, oc = pn.run_batch(rows, return_outcomes=True)
                kw = {k: v for k, v in noise.items() if k != "circuit_noise"}
                ref, roc = matrix_free.run_dm_batch(pat, rows, input_states=ins, window_size=w, noise=noise.get("circuit_noise"),
                                                    noise_kwargs=kw, return_outcomes=True)
                assert dm_distance(batch, ref) < 1e-10 and np.array_equal(oc, roc), (c["name"], w, noise)
        ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=c["window_size"])
        ps.reset()
        for node in ps.schedule_measure:  # step-by-step API
            a = ang[gs.trainable_nodes.index(node)] if node in gs.trainable_nodes else None
            st, o = ps.measure(a)
            assert o == c["outcomes"][str(node)]
        final = ps.reorder_qubits(st, ps.current_simulated_nodes(), gs.quantum_output_nodes)
        assert dm_distance(final, want) < 1e-10
    with pytest.raises(NotImplementedError):
        mb.PatternSimulator(gs, backend="cuda-dm", force0=False)


def test_dev_mode_matches_reference_golden():
    """dev_mode scheduling (np_simulator_sv.py:173-203, np_simulator_dm.py:160-201) on both CUDA
    backends against outputs recorded from the reference with wires of unequal length -- where the
    order, and the result, differ from the plain schedule --, plus batch == single run and the
    step-by-step API (window content after every measurement)."""
    for c in load_golden("dev_mode.json")["cases"]:
        name, args, kw = c["spec"]
        gs = getattr(mb.templates, name)(*args, **kw)
        inp = from_cplx(c["input_state"])
        ang = np.asarray(c["angles"])
        for backend, ref_name in (("cuda-sv", "numpy-sv"), ("cuda-dm", "numpy-dm")):
            want = from_cplx(c[ref_name]["output"])
            ps = mb.PatternSimulator(gs, input_state=inp, backend=backend, window_size=c["window_size"], dev_mode=True,
                                     wires=c["wires"])
            got = ps.run(ang)
            assert list(ps.outcomes.keys()) == c[ref_name]["order"]
            if backend == "cuda-sv":
                assert got.shape == want.shape and np.abs(got - want).max() < 1e-9
            else:
                assert dm_distance(got, want) < 1e-10
            rows = np.random.default_rng(c["seed"]).uniform(0, 2 * np.pi, (9, len(ang)))
            rows[4] = ang
            batch = ps.run_batch(rows) if backend == "cuda-dm" else ps.run_batch(rows, output_form="dm")  # run() defaults to 'dm'
            assert np.abs(batch[4] - got).max() < 1e-12
            ps.reset()
            for k, node in enumerate(ps.schedule_measure):
                assert node in ps.current_simulated_nodes()
                a = ang[gs.trainable_nodes.index(node)] if node in gs.trainable_nodes else None
                st = ps.measure(a)[0]
                assert node not in ps.current_simulated_nodes()
            target = gs.quantum_output_nodes if backend == "cuda-dm" else ps.simulator.plan.output_nodes
            final = ps.reorder_qubits(st, ps.current_simulated_nodes(), target)
            if backend == "cuda-sv":
                final = np.outer(final, final.conj())
            assert np.abs(final - got).max() < 1e-9
            if c[ref_name]["differs_from_plain_schedule"]:
                plain = mb.PatternSimulator(gs, input_state=inp, backend=backend, window_size=c["window_size"]).run(ang)
                assert np.abs(plain - got).max() > 1e-6


def test_expressivity_samples_through_the_batched_simulator():
    """mentpy_b200.tooling.expressivity (the role of utils/expressivity.py:39-125): the fidelity samples
    are one batched launch; checked against the oracle run on the same seeded inputs and angles, on
    both backends; a deep grid pattern comes out close to Haar."""
    from mentpy_b200 import tooling as tl
    from mentpy_b200.utils import generate_haar_random_states

    gs = mb.templates.grid_cluster(2, 6)
    n = 300
    f_sv = tl.sample_probability_density_of_fidelities(gs, n_samples=n, backend="cuda-sv", seed=3)
    f_dm = tl.sample_probability_density_of_fidelities(gs, n_samples=n, backend="cuda-dm", seed=3)
    assert f_sv.shape == (n,) and np.all((f_sv > -1e-12) & (f_sv < 1 + 1e-12)) and np.abs(f_sv - f_dm).max() < 1e-10
    rng = np.random.default_rng(3)  # the same draws as inside the sampler
    states = np.asarray(generate_haar_random_states(2, n, int(rng.integers(1 << 31))))
    angles = rng.uniform(0, 2 * np.pi, (n, len(gs.trainable_nodes)))
    want = matrix_free.run_sv_batch(PatternData.from_circuit(gs), angles, input_states=states)
    assert np.abs(f_sv - np.abs(np.einsum("bi,bi->b", states.conj(), want)) ** 2).max() < 1e-10
    deep = tl.expressivity_with_histogram(mb.templates.grid_cluster(2, 9), n_samples=20000, n_bins=50, seed=1)
    frozen = tl.expressivity_with_histogram(gs, n_bins=50, samples=np.full(1000, 0.999))  # a pattern acting as the identity
    assert 0 <= deep < 0.05 < frozen
    assert mb.utils.dim_su(4) == 15 and mb.PauliOp("XZ").txt == "XZ"


def test_full_size_c3_c4_against_the_oracle():
    """BASELINE configs 3 and 4 at FULL size (4,096 DM angle sets with and without depolarizing noise;
    2^20 gradient vectors) compared with the numpy oracle on strided subsamples of the very batch the
    kernels ran -- every kernel path the batch size selects (specialised kernels), every CTA region."""
    from mentpy_b200.gradients import psr_gradient_batched

    gs = mb.templates.grid_cluster(3, 8)
    pat = PatternData.from_circuit(gs)
    ang = np.random.default_rng(12).uniform(0, 2 * np.pi, (4096, 21))
    idx = np.arange(5, 4096, 131)
    for kw, okw in (({}, {}), ({"circuit_noise": "depolarizing", "p": 0.01}, {"noise": "depolarizing", "noise_kwargs": {"p": 0.01}})):
        rho, oc = mb.PatternSimulator(gs, backend="cuda-dm", **kw).run_batch(ang, return_outcomes=True)
        want, woc = matrix_free.run_dm_batch(pat, ang[idx], return_outcomes=True, **okw)
        assert dm_distance(rho[idx], want) < 1e-10 and np.array_equal(oc[idx], woc)

    gs = mb.templates.grid_cluster(4, 5)
    pat = PatternData.from_circuit(gs)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    B = 1 << 20
    X = torch.rand((B, 16), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(8)) * (2 * np.pi)
    tgt = np.exp(1j * np.arange(16)) / 4.0
    g, c = psr_gradient_batched(ps, X, tgt, return_cost=True)
    assert g.shape == (B, 16) and bool(torch.isfinite(g).all())
    idx = np.array([0, 127, 128, 65_537, 524_288, 1_000_003, B - 1])
    xs = X[torch.from_numpy(idx).cuda()].cpu().numpy()
    cost = lambda a: 1 - np.abs(matrix_free.run_sv_batch(pat, a) @ tgt.conj()) ** 2  # noqa: E731
    assert np.abs(c[torch.from_numpy(idx).cuda()].cpu().numpy() - cost(xs)).max() < 1e-11
    for i in range(16):
        e = np.zeros(16)
        e[i] = 1.5
        want = (cost(xs + e) - cost(xs - e)) / 3.0
        assert np.abs(g[torch.from_numpy(idx).cuda(), i].cpu().numpy() - want).max() < 1e-10
