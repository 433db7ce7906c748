"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.json from the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    python -m oracle.gen_golden

Everything written here comes from importing bestquark/mentpy as-is (through oracle/ref_shim.py)
and calling its public API: templates, MBQCircuit attributes, PatternSimulator(numpy-sv|numpy-dm),
mentpy.gradients.get_gradient, mentpy.optimizers.*.  The fixtures are what pins the oracles
(oracle/dense_port.py, oracle/matrix_free.py) and, through them and directly, the CUDA path.
Angles follow SURVEY.md section 8d: np.random.default_rng(seed).uniform(0, 2*pi, T).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_shim import import_reference  # noqa: E402
from oracle.pattern_data import PatternData  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def cplx(a):
    a = np.asarray(a, dtype=complex)
    return {"shape": list(a.shape), "re": a.real.reshape(-1).tolist(), "im": a.imag.reshape(-1).tolist()}


def template_specs():
    specs = []
    for n in range(2, 10):
        specs.append(("linear_cluster", [n], {}))
    for r in range(1, 5):
        for c in range(2, 9):
            if r * c <= 28:
                specs.append(("grid_cluster", [r, c], {}))
    specs.append(("grid_cluster", [3, 4], {"periodic": True}))
    specs.append(("grid_cluster", [4, 5], {}))
    for wires in ([2, 3, 4], [3, 3], [5, 2, 2], [4]):
        specs.append(("many_wires", [wires], {}))
    specs.append(("muta", [2, 1], {}))
    specs.append(("muta", [2, 1], {"one_column": True}))
    specs.append(("muta", [2, 2], {"one_column": True}))
    specs.append(("muta", [3, 1], {"one_column": True}))
    specs.append(("muta", [3, 1], {}))
    specs.append(("spturb", [4, 1], {}))
    specs.append(("spturb", [5, 1], {"periodic": True}))
    specs.append(("spturb", [4, 2], {}))
    lin = lambda n: ["linear_cluster", [n], {}]
    grid = lambda r, c: ["grid_cluster", [r, c], {}]
    specs.append(("vstack", [lin(3), lin(4)], {}))
    specs.append(("vstack", [grid(2, 3), lin(3), lin(2)], {}))
    specs.append(("hstack", [grid(2, 3), grid(2, 2)], {}))
    specs.append(("hstack", [lin(3), lin(2), lin(4)], {}))
    specs.append(("merge", [grid(2, 3), lin(4), [[2, 0]]], {}))
    specs.append(("merge", [grid(2, 3), grid(2, 2), [[5, 0], [2, 2]]], {}))
    seen, out = set(), []
    for s in specs:
        key = json.dumps(s)
        if key not in seen:
            seen.add(key)
            out.append(s)
    return out


def build(mp, spec):
    name, args, kwargs = spec
    if name in ("vstack", "hstack"):   # composite: args = list of sub-specs
        return getattr(mp, name)([build(mp, tuple(a)) for a in args])
    if name == "merge":               # args = [spec_a, spec_b, along]
        return mp.merge(build(mp, tuple(args[0])), build(mp, tuple(args[1])), along=[tuple(x) for x in args[2]])
    return getattr(mp.templates, name)(*args, **kwargs)


def structure_record(mp, spec):
    gs = build(mp, spec)
    pat = PatternData.from_circuit(gs)
    rec = {"spec": spec, "pattern": pat.to_json()}
    rec["nodes"] = [int(v) for v in gs.graph.nodes()]
    rec["flow"] = {str(v): int(gs.flow(v)) for v in gs.outputc} if gs.flow is not None else None
    rec["layers"] = [[int(v) for v in layer] for layer in gs.gflow.layers] if gs.gflow.layers else None
    rec["depth"] = int(gs.depth) if gs.flow is not None else None
    rec["planes"] = {str(k): v for k, v in gs.planes.items()}
    rec["outputc"] = [int(v) for v in gs.outputc]
    rec["inputc"] = [int(v) for v in gs.inputc]
    return rec


def haar_state(nq, seed):
    from scipy.stats import unitary_group

    return unitary_group.rvs(2**nq, random_state=seed)[:, 0]


def sim_case(mp, spec, backend, seed, window_size=None, x_nodes=(), fixed=None, haar=False,
             output_form="sv", trace=False):
    gs = build(mp, spec)
    for v in x_nodes:
        gs[v] = mp.Ment("X")
    for v, (ang, plane) in (fixed or {}).items():
        gs[int(v)] = mp.Ment(ang, plane)
    T = len(gs.trainable_nodes)
    angles = np.random.default_rng(seed).uniform(0, 2 * np.pi, T)
    kw = {}
    if window_size is not None:
        kw["window_size"] = window_size
    inp = haar_state(len(gs.input_nodes), seed) if haar else None
    ps = mp.PatternSimulator(gs, input_state=inp, backend=backend, **kw)
    rec = {
        "spec": spec, "backend": backend, "seed": seed, "window_size": int(ps.window_size),
        "x_nodes": [int(v) for v in x_nodes],
        "fixed": {str(k): [v[0], v[1]] for k, v in (fixed or {}).items()},
        "pattern": PatternData.from_circuit(gs).to_json(),
        "angles": angles.tolist(),
        "input_state": None if inp is None else cplx(inp),
    }
    if trace:
        steps = []
        ps.reset()
        for node in ps.schedule_measure:
            a = angles[gs.trainable_nodes.index(node)] if node in gs.trainable_nodes else gs[node].angle
            st, _ = ps.measure(a)
            steps.append(cplx(st))
        rec["trace"] = steps
        ps.reset()
    if backend == "numpy-sv":
        out = ps.run(angles, output_form=output_form)
        rec["output_form"] = output_form
    else:
        out = ps.run(angles)
        rec["output_form"] = "dm"
        rec["outcomes"] = {str(k): int(v) for k, v in ps.outcomes.items()}
    rec["output"] = cplx(out)
    return rec


def xyz_cases(mp):
    """2d. two-angle XYZ-plane measurements (ment.py:239-251) on the density-matrix backend: fixed
    (t1, t2) tuples -- the only form the reference can run, a trainable XYZ node receives one float
    and raises -- on measured inner nodes and on a measured output node."""
    import warnings

    cases = []
    for spec, seed, w, nodes, haar in (
            (("grid_cluster", [2, 4], {}), 40, None, {2: (0.3, 1.1), 5: (2.3, -0.4)}, False),
            (("grid_cluster", [3, 4], {}), 41, None, {4: (1.0, 0.5), 6: (-2.0, 2.2)}, True),
            (("linear_cluster", [6], {}), 42, 3, {1: (0.7, 0.2), 4: (4.0, 1.3)}, True),
            (("grid_cluster", [2, 5], {}), 43, 5, {4: (0.4, 0.9), 3: (1.5, -1.0)}, True),
            (("grid_cluster", [2, 4], {}), 44, 6, {1: (5.0, 0.6), 6: (0.1, -0.3)}, True)):
        gs = build(mp, spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for v, ang in nodes.items():
                gs[v] = mp.Ment(ang, "XYZ")
        T = len(gs.trainable_nodes)
        angles = np.random.default_rng(seed).uniform(0, 2 * np.pi, T)
        inp = haar_state(len(gs.input_nodes), seed) if haar else None
        kw = {} if w is None else {"window_size": w}
        ps = mp.PatternSimulator(gs, input_state=inp, backend="numpy-dm", **kw)
        out = ps.run(angles)
        cases.append({"spec": spec, "seed": seed, "window_size": int(ps.window_size),
                      "xyz": {str(k): list(v) for k, v in nodes.items()},
                      "pattern": PatternData.from_circuit(gs).to_json(), "angles": angles.tolist(),
                      "input_state": None if inp is None else cplx(inp), "output": cplx(out),
                      "outcomes": {str(k): int(v) for k, v in ps.outcomes.items()}})
    with open(os.path.join(GOLDEN, "dm_xyz_plane.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden.py (--only xyz)", "source": "bestquark/mentpy (unmodified)", "cases": cases}, f)


# the controlled measurements of the fixture below, rebuilt identically on the reference's classes (here)
# and on the product's (tests): name -> builder(package, ControlMent class) -> circuit
CONTROL_CASES = {
    "false_branch_x": lambda m, C: _with(m.templates.linear_cluster(5), lambda gs: gs.__setitem__(2, C(gs[0].outcome, None, "XY", 0, "X"))),
    "true_branch_trainable": lambda m, C: _with(m.templates.linear_cluster(6), lambda gs: gs.__setitem__(3, C(gs[1].outcome == 0, None, "XY", 0.4, "XY"))),
    "two_reads_xz": lambda m, C: _with(m.templates.grid_cluster(2, 4), lambda gs: gs.__setitem__(5, C(~gs[1].outcome + gs[4].outcome, None, "XZ", 0.3, "XY"))),
    "reordered_yz": lambda m, C: _with(m.templates.grid_cluster(2, 4), lambda gs: gs.__setitem__(2, C(gs[6].outcome == 0, None, "YZ", 0.9, "XY"))),
    "false_branch_xyz": lambda m, C: _with(m.templates.grid_cluster(3, 4), lambda gs: gs.__setitem__(6, C((gs[0].outcome + gs[4].outcome) == 1, None, "XZ", (0.5, 1.0), "XYZ"))),
    "two_controls": lambda m, C: _with(m.templates.grid_cluster(2, 5), lambda gs: (gs.__setitem__(2, C(gs[0].outcome, None, "XY", 1.1, "XY")),
                                                                                   gs.__setitem__(7, C((gs[5].outcome * gs[1].outcome) == 0, None, "YZ", 0.2, "XZ")))),
    "real_outcome_1": lambda m, C: _with(m.templates.many_wires([3, 3]), lambda gs: gs.__setitem__(1, C(gs[0].outcome, None, "XY", 0, "X"))),
}


def _with(gs, f):
    f(gs)
    return gs


def control_cases(mp):
    """2e. outcome-controlled measurements (operators/controlled_ment.py:14-113) on the density-matrix
    backend.  Under force0 outcomes are 0 unless prob0 < 1e-4, so the conditions are chosen to reach
    both branches; the last case re-uses the angles of the outcome-1 fixture, where node 0 really
    yields 1 and the condition fires on it."""
    import warnings

    from mentpy.operators import ControlMent

    quirk = json.load(open(os.path.join(GOLDEN, "dm_outcome_quirk.json")))
    cases = []
    for seed, (name, builder) in enumerate(CONTROL_CASES.items(), start=50):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gs = builder(mp, ControlMent)
        T = len(gs.trainable_nodes)
        angles = np.random.default_rng(seed).uniform(0, 2 * np.pi, T)
        inp, kw = None, {}
        if name == "real_outcome_1":
            angles = np.asarray(quirk["runs"][0]["angles"])
            q = quirk["input_state"]
            inp = (np.asarray(q["re"]) + 1j * np.asarray(q["im"])).reshape(q["shape"])
            kw = {"window_size": quirk["window_size"]}
        elif seed % 2:
            inp = haar_state(len(gs.input_nodes), seed)
        ps = mp.PatternSimulator(gs, input_state=inp, backend="numpy-dm", **kw)
        out = ps.run(angles)
        cases.append({"name": name, "seed": seed, "window_size": int(ps.window_size),
                      "measurement_order": [int(v) for v in gs.measurement_order],
                      "trainable_nodes": [int(v) for v in gs.trainable_nodes],
                      "pattern": PatternData.from_circuit(gs).to_json(), "angles": angles.tolist(),
                      "input_state": None if inp is None else cplx(inp), "output": cplx(out),
                      "outcomes": {str(k): int(v) for k, v in ps.outcomes.items()}})
    with open(os.path.join(GOLDEN, "dm_controlled.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden.py (--only ctl)", "source": "bestquark/mentpy (unmodified)", "cases": cases}, f)


def flow_wires(gs):
    """The wires of a pattern with causal flow: the chain input -> f(input) -> ... -> output."""
    wires = []
    for v in gs.input_nodes:
        wire = [int(v)]
        while wire[-1] not in gs.output_nodes:
            wire.append(int(gs.flow(wire[-1])))
        wires.append(wire)
    return wires


def dev_mode_cases(mp):
    """2f. dev_mode scheduling (np_simulator_sv.py:173-203, np_simulator_dm.py:160-201): the window
    measures the first node whose next wire-neighbour is present.  For wires of unequal length the
    order -- and the result -- differs from the plain schedule."""
    cases = []
    for spec, seed, w in ((("many_wires", [[3, 4, 2]], {}), 60, None), (("many_wires", [[3, 4, 2]], {}), 61, 5),
                          (("many_wires", [[2, 3, 3]], {}), 62, 4), (("grid_cluster", [2, 4], {}), 63, 3),
                          (("many_wires", [[4, 2]], {}), 64, None)):
        gs = build(mp, spec)
        wires = flow_wires(gs)
        T = len(gs.trainable_nodes)
        angles = np.random.default_rng(seed).uniform(0, 2 * np.pi, T)
        inp = haar_state(len(gs.input_nodes), seed)
        kw = {} if w is None else {"window_size": w}
        rec = {"spec": spec, "seed": seed, "wires": wires, "pattern": PatternData.from_circuit(gs).to_json(),
               "angles": angles.tolist(), "input_state": cplx(inp)}
        for backend in ("numpy-sv", "numpy-dm"):
            ps = mp.PatternSimulator(gs, input_state=inp, backend=backend, dev_mode=True, wires=wires, **kw)
            out = ps.run(angles)
            rec["window_size"] = int(ps.window_size)
            rec[backend] = {"output": cplx(out), "order": [int(k) for k in ps.outcomes.keys()]}
            plain = mp.PatternSimulator(gs, input_state=inp, backend=backend, **kw).run(angles)
            rec[backend]["differs_from_plain_schedule"] = bool(np.abs(plain - out).max() > 1e-6)
        cases.append(rec)
    with open(os.path.join(GOLDEN, "dev_mode.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden.py (--only dev)", "source": "bestquark/mentpy (unmodified)", "cases": cases}, f)


def main():
    mp = import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    if "--only" in sys.argv:
        {"xyz": xyz_cases, "ctl": control_cases, "dev": dev_mode_cases}[sys.argv[sys.argv.index("--only") + 1]](mp)
        return

    # 1. structure tables (integer indexing must be bit-exact)
    structures = [structure_record(mp, s) for s in template_specs()]
    with open(os.path.join(GOLDEN, "structures.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden.py", "source": "bestquark/mentpy (unmodified)",
                   "records": structures}, f)

    # 2. simulator known answers
    cases = []
    # BASELINE.json configs C1, C2, C4 (SV) and C3 (DM), seeds per SURVEY 8d
    cases.append(sim_case(mp, ("linear_cluster", [5], {}), "numpy-sv", 0, trace=True))
    cases.append(sim_case(mp, ("grid_cluster", [2, 6], {}), "numpy-sv", 1, trace=True))
    cases.append(sim_case(mp, ("grid_cluster", [2, 6], {}), "numpy-sv", 1, output_form="dm"))
    cases.append(sim_case(mp, ("grid_cluster", [3, 8], {}), "numpy-dm", 2))
    cases.append(sim_case(mp, ("grid_cluster", [4, 5], {}), "numpy-sv", 3))
    # Haar inputs, other windows, fixed-angle / X / Y nodes, muta (tests/test_simulators.py:33-64)
    cases.append(sim_case(mp, ("linear_cluster", [7], {}), "numpy-sv", 10, haar=True))
    cases.append(sim_case(mp, ("linear_cluster", [7], {}), "numpy-sv", 11, window_size=4, haar=True))
    cases.append(sim_case(mp, ("linear_cluster", [6], {}), "numpy-dm", 12, window_size=3, haar=True))
    cases.append(sim_case(mp, ("grid_cluster", [2, 5], {}), "numpy-sv", 13, window_size=5, x_nodes=(1, 7), haar=True))
    cases.append(sim_case(mp, ("grid_cluster", [2, 5], {}), "numpy-dm", 13, window_size=5, x_nodes=(1, 7), haar=True))
    cases.append(sim_case(mp, ("grid_cluster", [2, 4], {}), "numpy-sv", 14, window_size=4,
                          fixed={2: (0.3, "XY"), 5: (None, "Y")}))
    cases.append(sim_case(mp, ("grid_cluster", [2, 4], {}), "numpy-dm", 14, window_size=4,
                          fixed={2: (0.3, "XY"), 5: (None, "Y")}))
    cases.append(sim_case(mp, ("grid_cluster", [3, 5], {}), "numpy-sv", 15, haar=True, trace=True))
    cases.append(sim_case(mp, ("grid_cluster", [3, 4], {}), "numpy-dm", 16, haar=True))
    cases.append(sim_case(mp, ("grid_cluster", [3, 4], {"periodic": True}), "numpy-sv", 17, window_size=5))
    cases.append(sim_case(mp, ("many_wires", [[2, 3, 4]], {}), "numpy-sv", 18, haar=True))
    cases.append(sim_case(mp, ("many_wires", [[2, 3, 4]], {}), "numpy-dm", 18, haar=True))
    cases.append(sim_case(mp, ("muta", [2, 1], {}), "numpy-sv", 19, window_size=5, haar=True))
    cases.append(sim_case(mp, ("muta", [2, 1], {}), "numpy-dm", 19, window_size=5, haar=True))
    cases.append(sim_case(mp, ("muta", [2, 1], {"one_column": True}), "numpy-sv", 20, window_size=4))
    # too-small window: reference silently drops CZs -- reproducible quirk
    cases.append(sim_case(mp, ("muta", [2, 1], {"one_column": True}), "numpy-sv", 21))
    # DM planes beyond XY (deterministic under force0): XZ, YZ fixed + trainable
    cases.append(sim_case(mp, ("grid_cluster", [2, 4], {}), "numpy-dm", 22, window_size=4,
                          fixed={1: (0.7, "XZ"), 4: (None, "YZ")}))
    with open(os.path.join(GOLDEN, "sim_cases.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden.py", "source": "bestquark/mentpy (unmodified)",
                   "cases": cases}, f)

    # 2b. DM outcome-1 quirk (np_simulator_dm.py:335-338): window_size == |I| means the first
    # measurement hits the bare input; with input |-> x |+> and angle 0 on node 0, prob0 = 0.
    gs = mp.templates.many_wires([3, 3])
    minus = np.array([1.0, -1.0]) / np.sqrt(2)
    plus = np.array([1.0, 1.0]) / np.sqrt(2)
    inp = np.kron(minus, plus)
    ps = mp.PatternSimulator(gs, input_state=inp, backend="numpy-dm", window_size=2)
    quirk = []
    for ang in ([0.0, 0.4, 1.3, 2.2], [1e-3, 0.4, 1.3, 2.2], [0.5, 0.4, np.pi, 2.2]):
        ps.reset()
        out = ps.run(np.array(ang))
        quirk.append({"angles": ang, "output": cplx(out),
                      "outcomes": {str(k): int(v) for k, v in ps.outcomes.items()}})
    with open(os.path.join(GOLDEN, "dm_outcome_quirk.json"), "w") as f:
        json.dump({"spec": ("many_wires", [[3, 3]], {}), "window_size": 2,
                   "pattern": PatternData.from_circuit(gs).to_json(),
                   "input_state": cplx(inp), "runs": quirk}, f)

    # 2c. plane-Z measurements in mode="expectation" (np_simulator_dm.py:327-344): the qubit is traced
    # out unprojected and the recorded "outcome" is prob1 -- the classifier read-out of a pattern
    zcases = []
    for spec, seed, w, planes, haar in (
            (("grid_cluster", [2, 4], {}), 30, None, {3: "Z", 7: "Z"}, False),
            (("grid_cluster", [2, 5], {}), 31, 4, {2: "Z", 6: "X"}, True),
            (("linear_cluster", [6], {}), 32, 3, {2: "Z", 5: "Z"}, True),
            (("grid_cluster", [3, 4], {}), 33, None, {3: "Z", 7: "Z", 11: "Z", 5: "Y"}, True)):
        gs = build(mp, spec)
        for v, pl in planes.items():
            gs[v] = mp.Ment(pl)
        T = len(gs.trainable_nodes)
        angles = np.random.default_rng(seed).uniform(0, 2 * np.pi, T)
        inp = haar_state(len(gs.input_nodes), seed) if haar else None
        kw = {} if w is None else {"window_size": w}
        ps = mp.PatternSimulator(gs, input_state=inp, backend="numpy-dm", **kw)
        out = ps.run(angles, mode="expectation")
        zcases.append({"spec": spec, "seed": seed, "window_size": int(ps.window_size),
                       "planes": {str(k): v for k, v in planes.items()},
                       "pattern": PatternData.from_circuit(gs).to_json(), "angles": angles.tolist(),
                       "input_state": None if inp is None else cplx(inp), "output": cplx(out),
                       "outcomes": {str(k): float(v) for k, v in ps.outcomes.items()}})
    with open(os.path.join(GOLDEN, "dm_z_expectation.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden.py", "source": "bestquark/mentpy (unmodified)", "cases": zcases}, f)

    xyz_cases(mp)
    control_cases(mp)
    dev_mode_cases(mp)

    # 3. gradient + optimiser known answers: grid_cluster(4,5), cost 1 - <t|rho|t>
    gs = mp.templates.grid_cluster(4, 5)
    ps = mp.PatternSimulator(gs, backend="numpy-sv")
    tgt = np.full(16, 0.25)

    def cost(x):
        ps.reset()
        rho = ps.run(x)
        return float(1 - np.real(tgt.conj() @ rho @ tgt))

    x0 = np.random.default_rng(4).uniform(0, 2 * np.pi, 16)
    grad_psr = mp.gradients.get_gradient(cost, x0)
    grad_fd = mp.gradients.get_gradient(cost, x0, method="fd")
    rec = {"spec": ("grid_cluster", [4, 5], {}), "x": x0.tolist(), "target": cplx(tgt),
           "cost": cost(x0), "psr": grad_psr.tolist(), "fd": grad_fd.tolist()}

    gs_s = mp.templates.grid_cluster(2, 4)
    ps_s = mp.PatternSimulator(gs_s, backend="numpy-sv")
    tgt_s = haar_state(2, 5)

    def cost_s(x):
        ps_s.reset()
        rho = ps_s.run(x)
        return float(1 - np.real(tgt_s.conj() @ rho @ tgt_s))

    xs = np.random.default_rng(5).uniform(0, 2 * np.pi, len(gs_s.trainable_nodes))
    opt_rec = {"spec": ("grid_cluster", [2, 4], {}), "x": xs.tolist(), "target": cplx(tgt_s),
               "cost": cost_s(xs), "psr": mp.gradients.get_gradient(cost_s, xs).tolist()}
    opt_rec["hessian_psr_00_01"] = [float(v) for v in mp.gradients.get_hessian(cost_s, xs)[0, :2]]
    adam = mp.optimizers.AdamOptimizer(step_size=0.1)
    opt_rec["adam_5"] = adam.optimize(cost_s, xs.copy(), num_iters=5).tolist()
    sgd = mp.optimizers.SGDOptimizer(step_size=0.2, momentum=0.9)
    opt_rec["sgd_mom_5"] = sgd.optimize(cost_s, xs.copy(), num_iters=5).tolist()
    sgdn = mp.optimizers.SGDOptimizer(step_size=0.2, momentum=0.9, nesterov=True)
    opt_rec["sgd_nesterov_5"] = sgdn.optimize(cost_s, xs.copy(), num_iters=5).tolist()
    import random

    random.seed(7)
    rcd = mp.optimizers.RCDOptimizer(step_size=0.3, adaptive=True)
    opt_rec["rcd_seed7_6"] = rcd.optimize(cost_s, xs.copy(), num_iters=6).tolist()
    # data-set averaged training cost of docs/tutorials/intro-to-mbqml.rst:16-86 (same pattern:
    # muta(2, 1, one_column=True) with nodes 3 and 8 measured in X; numpy-sv instead of the absent
    # PennyLane backend; fidelity <t|rho|t> restated because calculator.fidelity is PennyLane's)
    gs_d = mp.templates.muta(2, 1, one_column=True)
    gs_d[3] = mp.Ment("X")
    gs_d[8] = mp.Ment("X")
    ps_d = mp.PatternSimulator(gs_d, backend="numpy-sv")
    from scipy.stats import unitary_group

    u1 = unitary_group.rvs(2, random_state=77)
    gate = np.kron(u1 / np.sqrt(np.linalg.det(u1)), np.eye(2))
    xs_d = [haar_state(2, 100 + i) for i in range(6)]
    ys_d = [gate @ v for v in xs_d]

    def cost_d(thetas):
        acc = 0.0
        for vin, vt in zip(xs_d, ys_d):
            ps_d.reset(input_state=vin)
            rho = ps_d(thetas)
            acc += 1 - float(np.real(vt.conj() @ rho @ vt))
        return acc / len(xs_d)

    xd = np.random.default_rng(6).uniform(0, 2 * np.pi, len(gs_d.trainable_nodes))
    data_rec = {"spec": ("muta", [2, 1], {"one_column": True}), "x_nodes": [3, 8], "x": xd.tolist(),
                "inputs": cplx(np.array(xs_d)), "targets": cplx(np.array(ys_d)), "cost": cost_d(xd),
                "psr": mp.gradients.get_gradient(cost_d, xd).tolist(),
                "fd": mp.gradients.get_gradient(cost_d, xd, method="fd").tolist(),
                "adam_4": mp.optimizers.AdamOptimizer(step_size=0.08).optimize(cost_d, xd.copy(), num_iters=4).tolist()}
    with open(os.path.join(GOLDEN, "gradients.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden.py", "c4": rec, "small": opt_rec, "dataset": data_rec}, f)

    # 4. helper known answers (calculator / Ment)
    helpers = {}
    st = haar_state(3, 31)
    helpers["sum_trace_pure"] = {"psi": cplx(st), "idx0": cplx(mp.calculator.partial_trace(st, [0])),
                                 "idx1": cplx(mp.calculator.partial_trace(st, [1]))}
    rho = np.outer(st, st.conj())
    helpers["trace_mixed"] = {"idx0": cplx(mp.calculator.partial_trace(rho, [0])),
                              "idx2": cplx(mp.calculator.partial_trace(rho, [2]))}
    helpers["ment"] = []
    for plane, ang in (("XY", 0.4), ("XZ", 1.2), ("YZ", -0.7), ("X", None), ("Y", None), ("Z", None)):
        m = mp.Ment(plane) if ang is None else mp.Ment(ang, plane)
        p0, p1 = m.get_povm()
        helpers["ment"].append({"plane": plane, "angle": ang, "matrix": cplx(m.matrix()),
                                "p0": cplx(p0), "p1": cplx(p1), "trainable": bool(mp.Ment(plane).is_trainable())})
    helpers["swap_sequence"] = []
    ps = mp.PatternSimulator(mp.templates.linear_cluster(3), backend="numpy-sv")
    for src, dst in (([3, 1, 2, 0], [0, 1, 2, 3]), ([5, 9], [9, 5]), ([0, 1, 2], [2, 0, 1])):
        helpers["swap_sequence"].append({"src": src, "dst": dst, "swaps": [list(s) for s in ps.find_swaps(src, dst)]})
    with open(os.path.join(GOLDEN, "helpers.json"), "w") as f:
        json.dump(helpers, f)
    print("golden fixtures written to", GOLDEN)
    for name in sorted(os.listdir(GOLDEN)):
        print(f"  {name}: {os.path.getsize(os.path.join(GOLDEN, name))} bytes")


if __name__ == "__main__":
    main()
