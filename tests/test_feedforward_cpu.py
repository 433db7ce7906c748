"""CPU: outcome sampling with flow corrections -- the oracle's two restatements against each other
(brute force with the corrections applied as gates, pennylane_simulator.py:138-153, vs the windowed
adapted-angle form the kernels use), the Philox generator against its published known answers, and
the product's host-side feed-forward masks against the oracle's tables."""
import itertools

import numpy as np
import pytest
from scipy.stats import unitary_group

import mentpy_b200 as mb
from conftest import build_spec, load_golden
from mentpy_b200.plan import correction_sources, feedforward, lower, window_is_valid
from oracle import feedforward as off
from oracle import matrix_free
from oracle.pattern_data import PatternData

RECORDS = load_golden("structures.json")["records"]


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kat:
        got = off.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert got.tolist() == list(want)
    u = off.uniform_for(12345, np.arange(200000), 3)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 2e-3
    assert not np.array_equal(u, off.uniform_for(12345, np.arange(200000), 4))


def _cases(max_nodes):
    for rec in RECORDS:
        pat = PatternData.from_json(rec["pattern"])
        if pat.n_nodes > max_nodes or rec["flow"] is None:
            continue
        if any(m is not None and m[0] not in ("XY", "X", "Y") for m in pat.measurements.values()):
            continue
        gs = build_spec(rec["spec"])
        n_meas = pat.n_nodes - len(pat.quantum_output_nodes)
        w = next((w for w in range(len(pat.input_nodes) + 1, min(n_meas, 6) + 1)
                  if window_is_valid(lower(gs, window_size=w))), None)
        if w is not None:
            yield rec, pat, gs, {int(k): v for k, v in rec["flow"].items()}, w


def test_windowed_adaptation_equals_physical_corrections():
    rng = np.random.default_rng(0)
    n = 0
    for rec, pat, gs, flow, w in _cases(9):
        T, M = len(pat.trainable_nodes), pat.n_nodes - len(pat.quantum_output_nodes)
        ang = rng.uniform(0, 2 * np.pi, T)
        inp = unitary_group.rvs(2 ** len(pat.input_nodes), random_state=3)[:, 0]
        det = matrix_free.run_sv_batch(pat, ang[None], input_states=inp[None], window_size=w)[0]
        recs = np.array(list(itertools.product((0, 1), repeat=M)))[:: max(1, 2**M // 8)]
        B = len(recs)
        out, oc, prob, _ = off.run_sv_sampled(pat, flow, np.repeat(ang[None], B, 0), window_size=w,
                                              input_states=np.repeat(inp[None], B, 0), forced=recs)
        assert np.array_equal(oc, recs)
        for b, r in enumerate(recs):
            p, rho = off.run_fullgraph_branch(pat, flow, ang, r, input_state=inp)
            assert abs(p - prob[b]) < 1e-12
            assert 1 - np.real(out[b].conj() @ rho @ out[b]) < 1e-12       # same branch state
            assert 1 - abs(np.vdot(det, out[b])) ** 2 < 1e-12              # = deterministic state
        n += 1
    assert n >= 20


def test_product_feedforward_masks_match_oracle_tables():
    n = 0
    for rec, pat, gs, flow, w in _cases(28):
        assert {v: gs.flow(v) for v in flow} == flow       # product flow == reference flow
        xs_o, zs_o = off.feedforward_tables(pat, flow)
        plan = lower(gs, window_size=w)
        xs, zs = correction_sources(gs, plan.schedule)
        assert {v: sorted(s) for v, s in xs.items()} == {v: sorted(s) for v, s in xs_o.items()}
        assert {v: sorted(s) for v, s in zs.items()} == {v: sorted(s) for v, s in zs_o.items()}
        ff = feedforward(gs, plan)
        step_of = {st.node: m for m, st in enumerate(plan.steps)}
        for m, st in enumerate(plan.steps):
            assert sorted(m - 1 - d for d in range(32) if ff[m].xdep >> d & 1) == sorted(step_of[i] for i in xs_o[st.node])
            assert sorted(m - 1 - d for d in range(32) if ff[m].zdep >> d & 1) == sorted(step_of[i] for i in zs_o[st.node])
        for q, v in enumerate(plan.output_nodes):
            assert sorted(m for m in range(len(ff)) if ff[m].outx >> q & 1) == sorted(step_of[i] for i in xs_o[v])
            assert sorted(m for m in range(len(ff)) if ff[m].outz >> q & 1) == sorted(step_of[i] for i in zs_o[v])
        n += 1
    assert n >= 40


def test_flow_adapt_angle():
    gs = mb.templates.grid_cluster(2, 3)
    fl = gs.gflow
    order = fl.measurement_order
    xs, zs = correction_sources(gs, order)
    outcomes = {0: 1, 3: 0, 1: 1}
    for v in (1, 4, 2):
        a = sum(outcomes.get(i, 0) for i in xs[v]) % 2
        b = sum(outcomes.get(i, 0) for i in zs[v]) % 2
        assert fl.adapt_angle(0.3, v, outcomes) == pytest.approx((-1) ** a * 0.3 + b * np.pi)
    measured = [v for v in order if v not in gs.output_nodes]
    assert len(fl.adapt_angles([0.1] * len(measured), outcomes)) == len(measured)

    class NoFlow:
        flow = None

    with pytest.raises(ValueError):
        correction_sources(NoFlow(), [0, 1])


def test_noisy_branch_average_is_trace_one():
    """Outcome-averaged noisy output (what the PennyLane backend returns) has trace 1 and is
    Hermitian; without noise it is the deterministic pure state."""
    rec = next(r for r in RECORDS if r["spec"][0] == "grid_cluster" and r["spec"][1] == [2, 3])
    pat = PatternData.from_json(rec["pattern"])
    flow = {int(k): v for k, v in rec["flow"].items()}
    ang = np.random.default_rng(1).uniform(0, 2 * np.pi, len(pat.trainable_nodes))
    avg, total = off.branch_average(pat, flow, ang, noise="amplitude_damping", noise_kwargs={"p": 0.2})
    assert abs(total - 1) < 1e-12 and abs(np.trace(avg) - 1) < 1e-12 and np.allclose(avg, avg.conj().T)
    avg0, _ = off.branch_average(pat, flow, ang)
    det = matrix_free.run_sv_batch(pat, ang[None])[0]
    assert np.allclose(avg0, np.outer(det, det.conj()), atol=1e-12)
