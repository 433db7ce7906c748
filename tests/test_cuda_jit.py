"""GPU: the three state-vector kernels behind mbqc_run_batch_sv -- run-time specialised
(sv_jit_src.inc), lean (sv_lean.cuh) and general register kernel (sv_reg.cuh) -- against the
reference's golden vectors and against each other on seeded batches (ragged sizes, Haar inputs,
fixed-angle nodes, both output forms).  Tolerance: infidelity <= 1e-10, amplitudes 1e-9."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import mentpy_b200 as mb
from conftest import ROOT, from_cplx, infidelity_pure, load_golden
from mentpy_b200 import _lib
from oracle import matrix_free
from oracle.pattern_data import PatternData

pytestmark = pytest.mark.gpu

CASES = [c for c in load_golden("sim_cases.json")["cases"] if c["backend"] == "numpy-sv" and c["window_size"] <= 5]


def _circuit(case):
    name, args, kwargs = case["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    for v in case["x_nodes"]:
        gs[v] = mb.Ment("X")
    for v, (ang, plane) in case["fixed"].items():
        gs[int(v)] = mb.Ment(ang, plane)
    return gs


@pytest.fixture
def jit_forced():
    lib = _lib.load()
    prev = lib.mbqc_jit_set_mode(2)
    yield lib
    lib.mbqc_jit_set_mode(prev)


@pytest.fixture
def jit_off():
    lib = _lib.load()
    prev = lib.mbqc_jit_set_mode(0)
    yield lib
    lib.mbqc_jit_set_mode(prev)


def _launched(lib):
    return lib.mbqc_jit_info().decode()


@pytest.mark.parametrize("case", CASES, ids=[f"{c['spec'][0]}{c['spec'][1]}-w{c['window_size']}-s{c['seed']}-{c['output_form']}" for c in CASES])
def test_specialised_kernel_reproduces_reference_goldens(case, jit_forced):
    gs = _circuit(case)
    inp = from_cplx(case["input_state"])
    ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv", window_size=case["window_size"])
    ang = np.asarray(case["angles"])
    want = from_cplx(case["output"])
    B = 67  # ragged: two full warps + 3 rows; row 5 carries the golden angles
    rows = np.random.default_rng(case["seed"]).uniform(0, 2 * np.pi, (B, len(ang)))
    rows[5] = ang
    before = _launched(jit_forced)
    got = ps.run_batch(torch.from_numpy(rows).cuda(), output_form=case["output_form"]).cpu().numpy()
    after = _launched(jit_forced)
    assert "failures=0" in after, after
    assert before != after or "compiled=0" not in after  # a specialised kernel was built or re-used
    if case["output_form"] == "sv":
        assert infidelity_pure(got[5], want) < 1e-10
        assert np.allclose(got[5], want, atol=1e-9, rtol=0)
    else:
        assert np.abs(got[5] - want).max() < 1e-10
    # the whole batch against the general kernels
    jit_forced.mbqc_jit_set_mode(0)
    ref = ps.run_batch(torch.from_numpy(rows).cuda(), output_form=case["output_form"]).cpu().numpy()
    jit_forced.mbqc_jit_set_mode(2)
    assert np.abs(got - ref).max() < 1e-12


@pytest.mark.parametrize("spec", [("linear_cluster", [5]), ("grid_cluster", [2, 6]), ("grid_cluster", [3, 5]),
                                  ("grid_cluster", [4, 5]), ("linear_cluster", [40]), ("muta", [2, 1])])
@pytest.mark.parametrize("B", [1, 31, 128, 4099])
def test_three_kernels_agree_with_the_oracle(spec, B, jit_forced):
    gs = getattr(mb.templates, spec[0])(*spec[1])
    T = len(gs.trainable_nodes)
    rng = np.random.default_rng(B + T)
    ang = rng.uniform(-7, 7, (B, T))
    want = matrix_free.run_sv_batch(PatternData.from_circuit(gs), ang)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    a = torch.from_numpy(ang).cuda()
    outs = {}
    jit_forced.mbqc_jit_set_mode(2)
    outs["jit"] = ps.run_batch(a).cpu().numpy()
    jit_forced.mbqc_jit_set_mode(0)
    outs["lean"] = ps.run_batch(a).cpu().numpy()
    assert "failures=0" in _launched(jit_forced)
    for name, got in outs.items():
        infid = np.abs(1 - np.abs(np.sum(got.conj() * want, axis=1)) ** 2)
        assert infid.max() < 1e-10, (name, infid.max())
    assert np.abs(outs["jit"] - outs["lean"]).max() < 1e-12


def test_general_register_kernel_still_matches(jit_off):
    """MBQC_SV_KERNEL_REG=1 routes everything through sv_reg_kernel (read once per process)."""
    code = (
        "import numpy as np, torch, mentpy_b200 as mb\n"
        "gs = mb.templates.grid_cluster(2, 6)\n"
        "ang = np.random.default_rng(3).uniform(0, 6.28, (1000, 10))\n"
        "ps = mb.PatternSimulator(gs, backend='cuda-sv')\n"
        "np.save('/tmp/_mbqc_reg_out.npy', ps.run_batch(torch.from_numpy(ang).cuda()).cpu().numpy())\n")
    env = dict(os.environ, MBQC_SV_KERNEL_REG="1", MBQC_JIT="0", PYTHONPATH=ROOT)
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=ROOT, timeout=300)
    reg = np.load("/tmp/_mbqc_reg_out.npy")
    gs = mb.templates.grid_cluster(2, 6)
    ang = np.random.default_rng(3).uniform(0, 6.28, (1000, 10))
    lean = mb.PatternSimulator(gs, backend="cuda-sv").run_batch(torch.from_numpy(ang).cuda()).cpu().numpy()
    assert np.abs(reg - lean).max() < 1e-12


def test_inputs_per_sample_and_shared_and_dm_form(jit_forced):
    from scipy.stats import unitary_group

    gs = mb.templates.grid_cluster(2, 5)
    T = len(gs.trainable_nodes)
    B = 300
    rng = np.random.default_rng(11)
    ang = rng.uniform(0, 2 * np.pi, (B, T))
    ins = np.stack([unitary_group.rvs(4, random_state=s)[:, 0] for s in range(B)])
    pat = PatternData.from_circuit(gs)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    got = ps.run_batch(torch.from_numpy(ang).cuda(), input_states=torch.from_numpy(ins).cuda()).cpu().numpy()
    want = matrix_free.run_sv_batch(pat, ang, input_states=ins)
    assert np.abs(1 - np.abs(np.sum(got.conj() * want, axis=1)) ** 2).max() < 1e-10
    assert np.allclose(got, want, atol=1e-9, rtol=0)
    ps1 = mb.PatternSimulator(gs, input_state=ins[3], backend="cuda-sv")
    got1 = ps1.run_batch(torch.from_numpy(ang).cuda(), output_form="dm").cpu().numpy()
    want1 = matrix_free.run_sv_batch(pat, ang, input_states=np.tile(ins[3], (B, 1)))
    assert np.abs(got1 - want1[:, :, None] * want1.conj()[:, None, :]).max() < 1e-10


def test_out_of_range_angles_are_reported(jit_forced):
    gs = mb.templates.grid_cluster(2, 6)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    ang = np.random.default_rng(0).uniform(0, 2 * np.pi, (64, 10))
    big = ang.copy()
    big[:, 3] += 2 * np.pi * 1e5  # far outside one turn, still inside the kernels' range (|x| < 2^31)
    a = ps.run_batch(torch.from_numpy(ang).cuda()).cpu().numpy()
    b = ps.run_batch(torch.from_numpy(big).cuda()).cpu().numpy()
    assert np.abs(a - b).max() < 1e-8  # the angle itself carries ~1e-10 of rounding at 6e5
    bad = ang.copy()
    bad[7, 2] = 3e9
    bad[9, 0] = np.inf
    ps.run_batch(torch.from_numpy(bad).cuda())
    st = ps.simulator.last_status.cpu().numpy()
    assert st[7] == _lib.STATUS_BAD_NORM and st[9] == _lib.STATUS_BAD_NORM and st[[0, 1, 8, 10]].max() == 0


# ---- specialised gradient kernel (sv_jit_grad_src.inc) -------------------------------------------
GRAD = load_golden("gradients.json")


def test_specialised_gradient_reproduces_the_reference_golden(jit_forced):
    """grid_cluster(4,5): cost and parameter-shift gradient recorded from mentpy.gradients.get_gradient."""
    from mentpy_b200.gradients import psr_gradient_batched

    g = GRAD["c4"]
    name, args, kwargs = g["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    x = np.asarray(g["x"])
    tgt = from_cplx(g["target"]).reshape(-1)
    rows = np.random.default_rng(0).uniform(0, 2 * np.pi, (37, len(x)))
    rows[11] = x
    grad, cost = psr_gradient_batched(ps, torch.from_numpy(rows).cuda(), tgt, return_cost=True)
    assert "failures=0" in jit_forced.mbqc_jit_info().decode()
    assert abs(cost[11].item() - g["cost"]) < 1e-12
    assert np.abs(grad[11].cpu().numpy() - np.asarray(g["psr"])).max() < 1e-12
    fd = psr_gradient_batched(ps, torch.from_numpy(rows).cuda(), tgt, shift=1e-5)
    assert np.abs(fd[11].cpu().numpy() - np.asarray(g["fd"])).max() < 1e-8  # central difference, h = 1e-5


@pytest.mark.parametrize("spec,w", [(("linear_cluster", [5]), None), (("grid_cluster", [2, 6]), None),
                                    (("grid_cluster", [3, 5]), None), (("grid_cluster", [4, 5]), None),
                                    (("grid_cluster", [2, 5]), 5), (("muta", [2, 1]), None),
                                    (("linear_cluster", [40]), 3), (("many_wires", [[3, 4, 2]]), None)])
def test_specialised_gradient_matches_general_kernels(spec, w, jit_forced):
    """Random targets, Haar inputs per sample, fixed-angle nodes: specialised kernel vs the
    ahead-of-time gradient kernels (which the round-1 suite pins to the reference)."""
    from scipy.stats import unitary_group

    from mentpy_b200.gradients import psr_gradient_batched

    gs = getattr(mb.templates, spec[0])(*spec[1])
    if spec[0] == "grid_cluster" and spec[1] == [2, 5]:
        gs[1] = mb.Ment("X")
        gs[7] = mb.Ment(0.3, "XY")
    kw = {} if w is None else {"window_size": w}
    ps = mb.PatternSimulator(gs, backend="cuda-sv", **kw)
    T, k, n_in = len(gs.trainable_nodes), len(gs.output_nodes), len(gs.input_nodes)
    B = 77
    rng = np.random.default_rng(T)
    ang = torch.from_numpy(rng.uniform(-4, 4, (B, T))).cuda()
    tgt = unitary_group.rvs(2**k, random_state=1)[:, 0] if k > 0 else np.ones(1, complex)
    ins = np.stack([unitary_group.rvs(2**n_in, random_state=s)[:, 0] for s in range(B)])
    res = {}
    for mode in (2, 0):
        jit_forced.mbqc_jit_set_mode(mode)
        g1, c1 = psr_gradient_batched(ps, ang, tgt, return_cost=True)
        g2, c2 = psr_gradient_batched(ps, ang, tgt, input_states=ins, return_cost=True)
        res[mode] = [x.cpu().numpy() for x in (g1, c1, g2, c2)]
    jit_forced.mbqc_jit_set_mode(2)
    assert "failures=0" in jit_forced.mbqc_jit_info().decode()
    for a, b in zip(res[2], res[0]):
        assert np.abs(a - b).max() < 1e-11


def test_specialised_gradient_dataset_mode(jit_forced):
    """Data-set averaged gradient (mbqc_psr_grad_dataset): sample p * S + s uses angle row p,
    input / target s."""
    from scipy.stats import unitary_group

    from mentpy_b200.gradients import psr_gradient_dataset

    gs = mb.templates.grid_cluster(2, 5)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    T = len(gs.trainable_nodes)
    P, S = 5, 9
    rng = np.random.default_rng(5)
    X = torch.from_numpy(rng.uniform(0, 2 * np.pi, (P, T))).cuda()
    ins = np.stack([unitary_group.rvs(4, random_state=s)[:, 0] for s in range(S)])
    tgs = np.stack([unitary_group.rvs(4, random_state=100 + s)[:, 0] for s in range(S)])
    out = {}
    for mode in (2, 0):
        jit_forced.mbqc_jit_set_mode(mode)
        g, c = psr_gradient_dataset(ps, X, tgs, ins, return_cost=True)
        out[mode] = (g.cpu().numpy(), c.cpu().numpy())
    jit_forced.mbqc_jit_set_mode(2)
    assert np.abs(out[2][0] - out[0][0]).max() < 1e-12 and np.abs(out[2][1] - out[0][1]).max() < 1e-12


# ---- regressions for the round-1 advisor findings -------------------------------------------------
def test_dataset_gradient_uses_the_simulators_input_state():
    """psr_gradient_dataset(input_states=None) must differentiate the cost that run_batch evaluates,
    i.e. use the simulator's own (non-default) input state."""
    from scipy.stats import unitary_group

    from mentpy_b200.gradients import psr_gradient_dataset

    gs = mb.templates.grid_cluster(2, 4)
    st = unitary_group.rvs(4, random_state=3)[:, 0]
    ps = mb.PatternSimulator(gs, input_state=st, backend="cuda-sv")
    T = len(gs.trainable_nodes)
    X = torch.from_numpy(np.random.default_rng(1).uniform(0, 2 * np.pi, (3, T))).cuda()
    tgs = np.stack([unitary_group.rvs(4, random_state=20 + s)[:, 0] for s in range(5)])
    g_none, c_none = psr_gradient_dataset(ps, X, tgs, None, return_cost=True)
    g_expl, c_expl = psr_gradient_dataset(ps, X, tgs, np.tile(st, (5, 1)), return_cost=True)
    assert np.abs((g_none - g_expl).cpu().numpy()).max() < 1e-14
    # and the cost is the one run_batch gives
    out = ps.run_batch(X).cpu().numpy()
    want = np.array([[1 - abs(np.vdot(t, o)) ** 2 for t in tgs] for o in out]).mean(axis=1)
    assert np.abs(c_none.cpu().numpy() - want).max() < 1e-12
    assert np.abs(c_none.cpu().numpy() - c_expl.cpu().numpy()).max() < 1e-14


def test_async_host_calls_with_per_call_input_state_and_dropped_handles():
    from scipy.stats import unitary_group

    gs = mb.templates.grid_cluster(2, 6)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    ang = np.random.default_rng(2).uniform(0, 2 * np.pi, (3000, 10))
    for s in range(6):  # a fresh temporary input tensor per call: it must be complete before the pipeline reads it
        st = unitary_group.rvs(4, random_state=s)[:, 0]
        want = ps.run_batch(torch.from_numpy(ang).cuda(), input_states=torch.from_numpy(st).cuda()).cpu().numpy()
        got = ps.run_batch_async(ang, input_states=st).result()
        assert np.abs(got - want).max() < 1e-12
    for _ in range(7):  # handles dropped without result(): their tickets must come back
        ps.run_batch_async(ang)
    import gc

    gc.collect()
    assert ps.run_batch_async(ang).result().shape == (3000, 4)


# ---- specialised density-matrix kernel (dm_jit_src.inc) -------------------------------------------
DM_CASES = [c for c in load_golden("sim_cases.json")["cases"] if c["backend"] == "numpy-dm" and 2 <= c["window_size"] <= 5
            and not any(str(p) == "Z" for _, p in c["fixed"].values())]


def _jit_counts(lib):
    info = lib.mbqc_jit_info().decode()
    return {k: int(v) for k, v in (kv.split("=") for kv in info.split() if kv.split("=")[0] in ("compiled", "from_disk", "failures"))}


@pytest.mark.parametrize("case", DM_CASES, ids=[f"{c['spec'][0]}{c['spec'][1]}-w{c['window_size']}-s{c['seed']}" for c in DM_CASES])
def test_specialised_dm_kernel_reproduces_reference_goldens(case, jit_forced):
    from conftest import dm_distance

    gs = _circuit(case)
    inp = from_cplx(case["input_state"])
    ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=case["window_size"])
    ang = np.asarray(case["angles"])
    B = 37  # ragged; row 5 carries the golden angles
    rows = np.random.default_rng(case["seed"]).uniform(0, 2 * np.pi, (B, len(ang)))
    rows[5] = ang
    got, oc = ps.run_batch(rows, return_outcomes=True)
    assert "failures=0" in _launched(jit_forced), _launched(jit_forced)
    assert dm_distance(got[5], from_cplx(case["output"])) < 1e-10
    jit_forced.mbqc_jit_set_mode(0)
    ref, roc = ps.run_batch(rows, return_outcomes=True)
    jit_forced.mbqc_jit_set_mode(2)
    assert np.abs(got - ref).max() < 1e-12
    assert np.array_equal(oc, roc)


@pytest.mark.parametrize("spec,w", [(("grid_cluster", [3, 8]), None), (("grid_cluster", [2, 5]), 5), (("grid_cluster", [4, 5]), None),
                                    (("linear_cluster", [6]), 3), (("linear_cluster", [7]), 2), (("linear_cluster", [5]), None),
                                    (("many_wires", [[3, 4, 2]]), None), (("grid_cluster", [2, 6]), None)])
@pytest.mark.parametrize("noise", [None, ("depolarizing", {"p": 0.01}), ("amplitude_damping", {"p": 0.2}),
                                   ("generalized_amplitude_damping", {"p": 0.1, "p_gad": 0.3})])
def test_specialised_dm_kernel_matches_oracle_and_general_kernel(spec, w, noise, jit_forced):
    from conftest import dm_distance
    from scipy.stats import unitary_group

    name, args = spec
    gs = getattr(mb.templates, name)(*args)
    pat = PatternData.from_circuit(gs)
    rng = np.random.default_rng(13)
    B, T = 261, len(gs.trainable_nodes)  # ragged: not a multiple of the samples per CTA
    ang = rng.uniform(0, 2 * np.pi, (B, T))
    kw = {} if w is None else {"window_size": w}
    nkw = {} if noise is None else {"circuit_noise": noise[0], **noise[1]}
    k = len(gs.input_nodes)
    inp = unitary_group.rvs(2**k, random_state=4)[:, 0] if k else None
    ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", **kw, **nkw)
    got, oc = ps.run_batch(ang, return_outcomes=True)
    after = _jit_counts(jit_forced)
    assert after["failures"] == 0, _launched(jit_forced)
    okw = None
    if noise is not None:
        okw = {"p": noise[1]["p"]}
        if "p_gad" in noise[1]:
            okw["p_gad"] = noise[1]["p_gad"]
    sub = slice(0, 9)
    want, woc = matrix_free.run_dm_batch(pat, ang[sub], window_size=(w or 1), input_states=inp, return_outcomes=True,
                                         **({} if noise is None else {"noise": noise[0], "noise_kwargs": okw}))
    assert dm_distance(got[sub], want) < 1e-10
    assert np.array_equal(oc[sub], woc)
    assert np.allclose(np.trace(got, axis1=1, axis2=2), 1.0, atol=1e-12)
    assert np.allclose(got, np.conj(np.swapaxes(got, 1, 2)), atol=1e-12)
    jit_forced.mbqc_jit_set_mode(0)
    ref, roc = ps.run_batch(ang, return_outcomes=True)
    jit_forced.mbqc_jit_set_mode(2)
    assert np.abs(got - ref).max() < 1e-12
    assert np.array_equal(oc, roc)


def test_specialised_dm_kernel_outcome1_rule(jit_forced):
    """np_simulator_dm.py:335-338: a step with prob0 < 1e-4 takes outcome 1.  The specialised kernel
    detects the step lazily and repeats the warp through its exact path; the rows next to the
    affected one must be untouched."""
    from conftest import dm_distance

    d = load_golden("dm_outcome_quirk.json")
    name, args, kwargs = d["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    ps = mb.PatternSimulator(gs, input_state=from_cplx(d["input_state"]), backend="cuda-dm", window_size=d["window_size"])
    T = len(gs.trainable_nodes)
    rng = np.random.default_rng(3)
    seen1 = False
    for run in d["runs"]:
        rows = rng.uniform(0, 2 * np.pi, (70, T))
        rows[33] = np.asarray(run["angles"])
        got, oc = ps.run_batch(rows, return_outcomes=True)
        assert dm_distance(got[33], from_cplx(run["output"])) < 1e-10
        want_oc = [run["outcomes"][str(n)] for n in ps.schedule_measure] if isinstance(run["outcomes"], dict) else None
        if want_oc is not None:
            assert list(oc[33]) == want_oc
            seen1 |= 1 in want_oc
        jit_forced.mbqc_jit_set_mode(0)
        ref, roc = ps.run_batch(rows, return_outcomes=True)
        jit_forced.mbqc_jit_set_mode(2)
        assert np.array_equal(oc, roc)
    # rows that divide by a small branch probability (product inputs behind dropped CZs reach prob0 ~ 1e-3,
    # hits subtract nearly equal blocks) amplify the kernels' different rounding: parity bar, not 1e-12
    assert np.abs(got - ref).max() < 1e-10
    assert seen1, "the golden file no longer exercises outcome 1"
    assert "failures=0" in _launched(jit_forced)


def test_specialised_dm_kernel_full_size_c3(jit_forced):
    """BASELINE configs[2] at full size (4,096 angle sets, grid_cluster(3,8), depolarizing 0.01):
    specialised vs general kernel on every row, trace / hermiticity / positivity invariants."""
    gs = mb.templates.grid_cluster(3, 8)
    ang = torch.from_numpy(np.random.default_rng(1).uniform(0, 2 * np.pi, (4096, len(gs.trainable_nodes)))).cuda()
    for nkw in ({}, {"circuit_noise": "depolarizing", "p": 0.01}):
        ps = mb.PatternSimulator(gs, backend="cuda-dm", **nkw)
        got = ps.run_batch(ang).cpu().numpy()
        jit_forced.mbqc_jit_set_mode(0)
        ref = ps.run_batch(ang).cpu().numpy()
        jit_forced.mbqc_jit_set_mode(2)
        assert np.abs(got - ref).max() < 1e-12
        assert np.allclose(np.trace(got, axis1=1, axis2=2), 1.0, atol=1e-12)
        assert np.linalg.eigvalsh(got[::64]).min() > -1e-12
    assert "failures=0" in _launched(jit_forced)


@pytest.mark.parametrize("wires,w,B", [([3, 3, 3], 3, 70), ([3, 3, 3, 3], 4, 70), ([3, 3, 3, 3], 4, 4200), ([3, 3, 3, 3, 3], 5, 70)])
def test_specialised_dm_kernel_exact_path_with_lanes(wires, w, B, jit_forced):
    """The lazily evaluated outcome rule on the multi-lane layouts (4 and 16 lanes per sample): with a
    window as small as the input register the first CZs are dropped (as in the reference), so an
    input qubit prepared orthogonal to its measurement direction gives prob0 = 0 -> outcome 1.  Rows
    with and without such a qubit share warps; compared with the oracle and the general kernel."""
    from conftest import dm_distance

    gs = mb.templates.many_wires(wires)
    pat = PatternData.from_circuit(gs)
    n_in, T = len(gs.input_nodes), len(gs.trainable_nodes)
    rng = np.random.default_rng(17)
    ang = rng.uniform(0, 2 * np.pi, (B, T))  # 4,200 rows at window 4: the 4-lane x 16-entry layout
    ins = np.zeros((B, 2**n_in), dtype=complex)
    hit = rng.random(B) < (0.3 if B < 1000 else 0.01)
    first = gs.trainable_nodes.index(gs.input_nodes[0])
    for b in range(B):
        qubits = []
        for q in range(n_in):
            v = rng.normal(size=2) + 1j * rng.normal(size=2)
            if q == 0 and hit[b]:  # orthogonal to (|0> + e^{i theta}|1>)/sqrt2 of the first measurement
                v = np.array([1.0, -np.exp(1j * ang[b, first])])
            qubits.append(v / np.linalg.norm(v))
        st = qubits[0]
        for v in qubits[1:]:
            st = np.kron(st, v)
        ins[b] = st
    ps = mb.PatternSimulator(gs, backend="cuda-dm", window_size=w)
    got, oc = ps.run_batch(ang, input_states=ins, return_outcomes=True)
    assert "failures=0" in _launched(jit_forced)
    want, woc = matrix_free.run_dm_batch(pat, ang, input_states=ins, window_size=w, return_outcomes=True)
    assert woc[:, 0].astype(bool).tolist() == hit.tolist()  # the construction does what it says
    assert np.array_equal(oc, woc) and dm_distance(got, want) < 1e-10
    assert np.array_equal((ps.last_status & 2) != 0, woc.any(axis=1))
    jit_forced.mbqc_jit_set_mode(0)
    ref, roc = ps.run_batch(ang, input_states=ins, return_outcomes=True)
    jit_forced.mbqc_jit_set_mode(2)
    assert np.array_equal(oc, roc)
    # rows that divide by a small branch probability (product inputs behind dropped CZs reach prob0 ~ 1e-3,
    # hits subtract nearly equal blocks) amplify the kernels' different rounding: parity bar, not 1e-12
    assert np.abs(got - ref).max() < 1e-10
