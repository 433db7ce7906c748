"""GPU: sampled runs (force0=False) -- Born-rule outcomes from Philox, flow-adapted angles,
byproduct corrections -- against the CPU oracle (oracle/feedforward.py) shot by shot, and through
size-independent properties at the benchmark sizes."""
import itertools

import numpy as np
import pytest
from scipy.stats import unitary_group

import mentpy_b200 as mb
from conftest import load_golden
from oracle import feedforward as off
from oracle import matrix_free
from oracle.pattern_data import PatternData

pytestmark = pytest.mark.gpu


def _flow_dict(gs):
    return {v: gs.flow(v) for v in gs.measurement_order if v not in gs.output_nodes}


def _fid(a, b):
    return np.abs(np.einsum("bi,bi->b", a.conj(), b)) ** 2


@pytest.mark.parametrize("name,args,w,fixed", [
    ("linear_cluster", [7], None, {}), ("linear_cluster", [6], 4, {2: "X"}), ("grid_cluster", [2, 5], None, {}),
    ("grid_cluster", [3, 4], 4, {1: "Y", 5: "X"}), ("grid_cluster", [4, 4], None, {}), ("grid_cluster", [2, 6], 5, {}),
    ("muta", [2, 1], 5, {}),
    # windows 6..12: the shared-memory sampled kernel (sv_smem_sample_kernel)
    ("grid_cluster", [2, 6], 6, {}), ("grid_cluster", [3, 5], 7, {3: "X"}), ("grid_cluster", [3, 4], 8, {}),
    ("linear_cluster", [12], 10, {}), ("grid_cluster", [4, 4], 12, {}),
])
def test_sampled_sv_matches_oracle_shot_by_shot(name, args, w, fixed):
    gs = getattr(mb.templates, name)(*args)
    for v, pl in fixed.items():
        gs[v] = mb.Ment(pl)
    kw = {} if w is None else {"window_size": w}
    ps = mb.PatternSimulator(gs, backend="cuda-sv", force0=False, seed=2024, **kw)
    pat = PatternData.from_circuit(gs)
    B, T, n_in = 257, len(gs.trainable_nodes), len(gs.input_nodes)
    rng = np.random.default_rng(3)
    ang = rng.uniform(0, 2 * np.pi, (B, T))
    ins = np.stack([unitary_group.rvs(2**n_in, random_state=s)[:, 0] for s in range(B)])
    got = ps.sample_batch(ang, input_states=ins, sample_offset=1000)
    want, oc, prob, (xb, zb) = off.run_sv_sampled(pat, _flow_dict(gs), ang, seed=2024, sample_offset=1000,
                                                  input_states=ins, window_size=ps.window_size)
    assert np.array_equal(got.outcomes, oc)                   # bit-exact outcome records
    assert np.array_equal(got.x, xb) and np.array_equal(got.z, zb)
    assert np.all(1 - _fid(got.states, want) < 1e-10)
    assert np.allclose(got.prob, prob, rtol=1e-10, atol=0)
    # corrected shots all equal the deterministic state
    det = mb.PatternSimulator(gs, backend="cuda-sv", **kw).run_batch(ang, input_states=ins)
    assert np.all(1 - _fid(got.states, det) < 1e-10)
    # uncorrected branch states + byproducts
    raw = ps.sample_batch(ang, input_states=ins, sample_offset=1000, correct=False)
    want_raw = off.run_sv_sampled(pat, _flow_dict(gs), ang, seed=2024, sample_offset=1000, input_states=ins,
                                  window_size=ps.window_size, correct=False)[0]
    assert np.array_equal(raw.outcomes, oc) and np.all(1 - _fid(raw.states, want_raw) < 1e-10)


def test_sampling_properties_at_benchmark_size():
    """C2 pattern, 65,536 shots: every corrected shot equals the force0 state; outcomes are fair
    coins for a cluster state; the stream is reproducible, splittable and seed dependent."""
    gs = mb.templates.grid_cluster(2, 6)
    B, T = 65536, len(gs.trainable_nodes)
    ang = np.random.default_rng(1).uniform(0, 2 * np.pi, (B, T))
    ps = mb.PatternSimulator(gs, backend="cuda-sv", force0=False, seed=7)
    res = ps.sample_batch(ang)
    det = mb.PatternSimulator(gs, backend="cuda-sv").run_batch(ang)
    assert np.all(1 - _fid(res.states, det) < 1e-10)
    freq = res.outcomes.mean(axis=0)
    assert np.all(np.abs(freq - 0.5) < 5 * 0.5 / np.sqrt(B))
    corr = np.corrcoef(res.outcomes.T.astype(float))
    assert np.abs(corr - np.eye(len(corr))).max() < 0.03
    assert np.allclose(res.prob, 2.0 ** -res.outcomes.shape[1], rtol=1e-9)
    again = mb.PatternSimulator(gs, backend="cuda-sv", force0=False, seed=7).sample_batch(ang)
    assert np.array_equal(again.outcomes, res.outcomes)
    halves = np.concatenate([ps.sample_batch(ang[:30000], sample_offset=0).outcomes,
                             ps.sample_batch(ang[30000:], sample_offset=30000).outcomes])
    assert np.array_equal(halves, res.outcomes)
    other = ps.sample_batch(ang, seed=8, sample_offset=0)
    assert not np.array_equal(other.outcomes, res.outcomes)
    nxt = ps.sample_batch(ang[:100])          # default offset continues after the first call
    assert not np.array_equal(nxt.outcomes, res.outcomes[:100])
    # CUDA tensors in -> CUDA tensors out
    import torch

    t = ps.sample_batch(torch.as_tensor(ang[:512]).cuda(), sample_offset=0)
    assert t.states.is_cuda and np.array_equal(t.outcomes.cpu().numpy(), res.outcomes[:512])


def test_forced_records_enumerate_all_branches():
    gs = mb.templates.grid_cluster(2, 4)
    gs[2] = mb.Ment("X")
    ps = mb.PatternSimulator(gs, backend="cuda-sv", force0=False)
    M, T = 6, len(gs.trainable_nodes)
    ang = np.random.default_rng(5).uniform(0, 2 * np.pi, T)
    inp = unitary_group.rvs(4, random_state=2)[:, 0]
    recs = np.array(list(itertools.product((0, 1), repeat=M)), dtype=np.int8)
    res = ps.sample_batch(np.repeat(ang[None], len(recs), 0), input_states=inp, forced_outcomes=recs)
    assert np.array_equal(res.outcomes, recs)
    assert abs(res.prob.sum() - 1) < 1e-12
    det = mb.PatternSimulator(gs, backend="cuda-sv", input_state=inp).run_batch(ang[None])[0]
    assert np.all(1 - np.abs(res.states @ det.conj()) ** 2 < 1e-10)
    pat = PatternData.from_circuit(gs)
    for b in (0, 21, 63):
        p, rho = off.run_fullgraph_branch(pat, _flow_dict(gs), ang, recs[b], input_state=inp)
        assert abs(p - res.prob[b]) < 1e-12 and 1 - np.real(res.states[b].conj() @ rho @ res.states[b]) < 1e-10


def test_sampled_dm_noiseless_equals_sv_and_run_api():
    gs = mb.templates.grid_cluster(3, 5)
    B, T = 300, len(gs.trainable_nodes)
    ang = np.random.default_rng(2).uniform(0, 2 * np.pi, (B, T))
    sv = mb.PatternSimulator(gs, backend="cuda-sv", force0=False, seed=11).sample_batch(ang)
    dm = mb.PatternSimulator(gs, backend="cuda-dm", force0=False, seed=11).sample_batch(ang)
    assert np.array_equal(sv.outcomes, dm.outcomes) and np.array_equal(sv.x, dm.x) and np.array_equal(sv.z, dm.z)
    assert np.allclose(dm.states, sv.states[:, :, None] * sv.states.conj()[:, None, :], atol=1e-11)
    assert np.allclose(dm.prob, sv.prob, rtol=1e-10)
    raw_sv = mb.PatternSimulator(gs, backend="cuda-sv", force0=False, seed=11).sample_batch(ang, correct=False)
    raw_dm = mb.PatternSimulator(gs, backend="cuda-dm", force0=False, seed=11).sample_batch(ang, correct=False)
    assert np.allclose(raw_dm.states, raw_sv.states[:, :, None] * raw_sv.states.conj()[:, None, :], atol=1e-11)
    # reference-style stateful API
    ps = mb.PatternSimulator(gs, backend="cuda-sv", force0=False, seed=3)
    rho = ps.run(ang[0])
    want = mb.PatternSimulator(gs, backend="cuda-sv").run(ang[0])
    assert np.allclose(rho, want, atol=1e-10)                        # 'dm' form: phase free
    assert set(ps.outcomes) == set(ps.schedule_measure) and set(ps.outcomes.values()) <= {0, 1}
    assert set(ps.byproducts) == set(gs.output_nodes)
    with pytest.raises(ValueError):
        ps.run(ang[0])                                               # needs reset, like the reference
    ps.reset()
    assert ps.run(ang[0], output_form="sv").shape == (8,)
    with pytest.raises(NotImplementedError):
        ps.reset() or ps.measure(0.1)


@pytest.mark.parametrize("noise", [{"circuit_noise": "depolarizing", "p": 0.1},
                                   {"circuit_noise": "amplitude_damping", "p": 0.25},
                                   {"circuit_noise": "phase_flip", "p": 0.2}])
def test_noisy_branches_average_to_the_pennylane_semantics(noise):
    """All 2^M outcome records forced through the noisy DM kernel: sum_b p_b = 1 and
    sum_b p_b rho_b (corrected) equals the brute-force full-graph simulation with the channels on
    every wire and the corrections applied as gates (pennylane_simulator.py:118-153)."""
    for shape, w in (([2, 3], 3), ([2, 4], 4)):
        gs = mb.templates.grid_cluster(*shape)
        pat = PatternData.from_circuit(gs)
        M, T = pat.n_nodes - 2, len(gs.trainable_nodes)
        ang = np.random.default_rng(8).uniform(0, 2 * np.pi, T)
        inp = unitary_group.rvs(4, random_state=4)[:, 0]
        recs = np.array(list(itertools.product((0, 1), repeat=M)), dtype=np.int8)
        ps = mb.PatternSimulator(gs, backend="cuda-dm", force0=False, window_size=w, **noise)
        res = ps.sample_batch(np.repeat(ang[None], len(recs), 0), input_states=inp, forced_outcomes=recs)
        assert abs(res.prob.sum() - 1) < 1e-12
        avg = np.einsum("b,bij->ij", res.prob, res.states)
        kw = {k: v for k, v in noise.items() if k != "circuit_noise"}
        want, total = off.branch_average(pat, _flow_dict(gs), ang, input_state=inp, noise=noise["circuit_noise"],
                                         noise_kwargs=kw)
        assert abs(total - 1) < 1e-12
        assert np.abs(avg - want).max() < 1e-12
        # the same through the public entry point (what the reference's PennyLane backend returns)
        assert np.abs(ps.simulator.outcome_averaged(ang, input_state=inp) - want).max() < 1e-12
        if shape == [2, 3]:  # without a channel the corrected branches coincide: the force0 state comes back
            clean = mb.PatternSimulator(gs, backend="cuda-sv", force0=False, window_size=w)
            pure = mb.PatternSimulator(gs, backend="cuda-dm", window_size=w).run_batch(ang[None], input_states=inp)[0]
            assert np.abs(clean.simulator.outcome_averaged(ang, input_state=inp) - pure).max() < 1e-10
        # sampled frequencies follow the branch probabilities (chi-square-ish on the first outcome)
        shots = ps.sample_batch(np.repeat(ang[None], 20000, 0), input_states=inp, seed=1)
        p_first1 = res.prob[recs[:, 0] == 1].sum()
        assert abs(shots.outcomes[:, 0].mean() - p_first1) < 5 * np.sqrt(p_first1 * (1 - p_first1) / 20000) + 1e-9
