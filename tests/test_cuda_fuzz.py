"""GPU: the CUDA path against the oracle on the seeded random patterns of fuzz_patterns.py (the
oracle is pinned against the live reference on the very same seeds by test_fuzz_reference_cpu.py)."""
import numpy as np
import pytest

import mentpy_b200 as mb
from conftest import dm_distance
from fuzz_patterns import random_pattern
from oracle import matrix_free
from oracle.pattern_data import PatternData

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(100))
def test_cuda_equals_oracle_on_random_patterns(seed):
    mixed = seed % 2 == 1
    gs, w, ang, inp = random_pattern(mb, seed, mixed)
    pat = PatternData.from_circuit(gs)
    rng = np.random.default_rng(1000 + seed)
    A = np.vstack([ang[None], rng.uniform(0, 2 * np.pi, (20, len(ang)))])
    if mixed:
        ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=w)
        got, oc = ps.run_batch(A, return_outcomes=True)
        want, woc = matrix_free.run_dm_batch(pat, A, input_states=inp[None], window_size=ps.window_size, return_outcomes=True)
        assert dm_distance(got, want) < 1e-10 and np.array_equal(oc, woc)
        rho = ps.run(ang)
        assert dm_distance(rho, want[0]) < 1e-10
    else:
        ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv", window_size=w)
        got = ps.run_batch(A)
        want = matrix_free.run_sv_batch(pat, A, input_states=inp[None], window_size=ps.window_size)
        assert np.max(1 - np.abs(np.sum(got.conj() * want, axis=1)) ** 2) < 1e-10
        assert np.allclose(got, want, atol=1e-9)           # including the reference's global phase
        # the streaming engine runs the same pattern (single angle set, state in HBM)
        st = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv-stream", window_size=w)
        psi = st.run(ang, output_form="sv")
        assert 1 - abs(np.vdot(psi, want[0])) ** 2 < 1e-10


@pytest.mark.parametrize("seed", range(0, 100, 2))
def test_sampled_runs_equal_oracle_on_random_patterns(seed):
    """force0=False on the random SV patterns: outcome records bit-exact, corrected states equal."""
    from oracle import feedforward as off

    gs, w, ang, inp = random_pattern(mb, seed, False)
    pat = PatternData.from_circuit(gs)
    flow = {v: gs.flow(v) for v in gs.measurement_order if v not in gs.output_nodes}
    A = np.vstack([ang[None], np.random.default_rng(2000 + seed).uniform(0, 2 * np.pi, (63, len(ang)))])
    if w > 12:  # windows 6..12 run on the shared-memory sampled kernel, beyond that sampled runs are refused
        with pytest.raises(NotImplementedError):
            mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv", window_size=w, force0=False)
        w = 12 if len(gs.input_nodes) <= 12 else None
        if w is None:
            pytest.skip("more inputs than the sampled kernels' window")
    ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv", window_size=w, force0=False, seed=seed)
    got = ps.sample_batch(A, sample_offset=5)
    # the oracle orders outputs like quantum_output_nodes; the plan may use output_nodes (SV rule)
    want, oc, prob, (xb, zb) = off.run_sv_sampled(pat, flow, A, seed=seed, sample_offset=5, input_states=inp[None],
                                                  window_size=ps.window_size)
    assert np.array_equal(got.outcomes, oc)
    assert np.allclose(got.prob, prob, rtol=1e-9)
    order = [pat.quantum_output_nodes.index(v) for v in ps.plan.output_nodes]
    assert np.array_equal(got.x, xb[:, order]) and np.array_equal(got.z, zb[:, order])
    k = len(order)
    want = want.reshape([len(A)] + [2] * k).transpose([0] + [1 + p for p in order]).reshape(len(A), -1)
    assert np.max(1 - np.abs(np.sum(got.states.conj() * want, axis=1)) ** 2) < 1e-10


@pytest.mark.parametrize("seed", range(1, 100, 4))
def test_noisy_dm_equals_oracle_on_random_patterns(seed):
    gs, w, ang, inp = random_pattern(mb, seed, True)
    pat = PatternData.from_circuit(gs)
    kinds = [("depolarizing", {"p": 0.07}), ("amplitude_damping", {"p": 0.15}), ("phase_damping", {"p": 0.2}),
             ("phase_flip", {"p": 0.1}), ("generalized_amplitude_damping", {"p": 0.1, "p_gad": 0.4})]
    kind, kw = kinds[seed % len(kinds)]
    A = np.vstack([ang[None], np.random.default_rng(3000 + seed).uniform(0, 2 * np.pi, (8, len(ang)))])
    ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=w, circuit_noise=kind, **kw)
    got = ps.run_batch(A)
    want = matrix_free.run_dm_batch(pat, A, input_states=inp[None], window_size=ps.window_size, noise=kind, noise_kwargs=kw)
    assert dm_distance(got, want) < 1e-10
    assert np.allclose(np.trace(got, axis1=1, axis2=2), 1.0, atol=1e-12)


@pytest.mark.parametrize("seed", range(0, 100, 4))
def test_gradients_equal_shifted_oracle_evaluations_on_random_patterns(seed):
    from mentpy_b200.gradients import psr_gradient_batched
    from scipy.stats import unitary_group

    gs, w, ang, inp = random_pattern(mb, seed, False)
    if w > 5:
        w = 5 if len(gs.input_nodes) < 5 else None
        if w is None:
            pytest.skip("window too large for the fused gradient")
    pat = PatternData.from_circuit(gs)
    ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv", window_size=w)
    k = len(gs.output_nodes)
    tgt = unitary_group.rvs(2**k, random_state=seed)[:, 0]
    X = np.vstack([ang[None], np.random.default_rng(4000 + seed).uniform(0, 2 * np.pi, (4, len(ang)))])
    g, c = psr_gradient_batched(ps, X, tgt, return_cost=True)

    def cost(Y):  # the oracle follows the same output-order rule as the plan
        psi = matrix_free.run_sv_batch(pat, Y, input_states=inp[None], window_size=ps.window_size)
        return 1 - np.abs(psi @ tgt.conj()) ** 2

    assert np.allclose(c, cost(X), atol=1e-11)
    T = len(ang)
    for i in range(T):
        e = np.zeros(T); e[i] = 1.5
        assert np.allclose(g[:, i], (cost(X + e) - cost(X - e)) / 3.0, atol=1e-10)


@pytest.mark.parametrize("seed", range(100, 140))
def test_user_schedules_equal_oracle(seed):
    """The `schedule=` kwarg on the CUDA backends (incl. streaming) on perturbed measurement orders;
    the oracle is pinned against the live reference on the same seeds (CPU suite)."""
    from fuzz_patterns import random_schedule

    mixed = seed % 2 == 1
    gs, w, ang, inp = random_pattern(mb, seed, mixed)
    sched = random_schedule(gs, seed, mixed)
    pat = PatternData.from_circuit(gs)
    A = np.vstack([ang[None], np.random.default_rng(5000 + seed).uniform(0, 2 * np.pi, (12, len(ang)))])
    if mixed:
        ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=w, schedule=sched)
        got, oc = ps.run_batch(A, return_outcomes=True)
        want, woc = matrix_free.run_dm_batch(pat, A, input_states=inp[None], window_size=w, schedule=sched, return_outcomes=True)
        assert dm_distance(got, want) < 1e-10 and np.array_equal(oc, woc)
    else:
        ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv", window_size=w, schedule=sched)
        got = ps.run_batch(A)
        want = matrix_free.run_sv_batch(pat, A, input_states=inp[None], window_size=w, schedule=sched)
        assert np.max(1 - np.abs(np.sum(got.conj() * want, axis=1)) ** 2) < 1e-10
        assert np.allclose(got, want, atol=1e-9)
        st = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv-stream", window_size=w, schedule=sched)
        assert 1 - abs(np.vdot(st.run(ang, output_form="sv"), want[0])) ** 2 < 1e-10
