"""CPU: host-side indexing (templates, relabelling, causal flow, layers, measurement order,
trainable order) must be bit-exact with tables dumped from the unmodified reference
(tests/golden/structures.json; mentpy/mbqc/mbqcircuit.py:325-422, flow.py:115-185)."""
import numpy as np
import pytest

import mentpy_b200 as mb
from conftest import build_spec, load_golden

RECORDS = load_golden("structures.json")["records"]


def _id(r):
    return f"{r['spec'][0]}{r['spec'][1]}{r['spec'][2] or ''}"


_build = build_spec


@pytest.mark.parametrize("rec", RECORDS, ids=[_id(r) for r in RECORDS])
def test_structure_tables(rec):
    gs = _build(rec["spec"])
    pat = rec["pattern"]
    assert list(gs.graph.nodes()) == rec["nodes"]
    assert sorted(tuple(sorted(e)) for e in gs.graph.edges()) == sorted(tuple(sorted(e)) for e in pat["edges"])
    assert gs.input_nodes == pat["input_nodes"]
    assert gs.output_nodes == pat["output_nodes"]
    assert gs.trainable_nodes == pat["trainable_nodes"]
    assert gs.measurement_order == pat["measurement_order"]
    assert gs.quantum_output_nodes == pat["quantum_output_nodes"]
    assert gs.outputc == rec["outputc"]
    assert gs.inputc == rec["inputc"]
    assert {str(k): v for k, v in gs.planes.items()} == rec["planes"]
    assert list(gs.measurements.keys()) == [int(k) for k in pat["measurements"].keys()]
    if rec["flow"] is None:   # no causal flow (spturb)
        assert gs.flow is None and gs.gflow.layers is None
    else:
        assert {str(v): gs.flow(v) for v in gs.outputc} == rec["flow"]
        assert gs.gflow.layers == rec["layers"]
        assert gs.depth == rec["depth"]
    assert len(gs) == pat["n_nodes"]


def test_config_table_survey_8():
    gs = mb.templates.grid_cluster(2, 6)
    assert gs.measurement_order == [0, 6, 1, 7, 2, 8, 3, 9, 4, 10, 5, 11]
    assert gs.trainable_nodes == [0, 1, 2, 3, 4, 6, 7, 8, 9, 10]
    gs = mb.templates.muta(2, 1)
    assert gs.measurement_order == [0, 4, 5, 1, 6, 2, 7, 3, 13, 8, 9, 14, 10, 15, 11, 16, 12, 17]
    assert (gs.input_nodes, gs.output_nodes) == ([0, 4], [12, 17])


def test_setitem_updates_trainables():
    gs = mb.templates.grid_cluster(2, 4)
    gs[1] = mb.Ment("X")
    gs[5] = mb.Ment(0.3, "XY")
    assert gs.trainable_nodes == [0, 2, 4, 6]
    assert gs[1].plane == "X" and gs[1].angle == 0 and gs[1].node_id == 1
    with pytest.raises(ValueError):
        gs[99] = mb.Ment("X")
    with pytest.raises(ValueError):
        gs[1] = "X"


def test_unsorted_labels_are_ranked():
    g = mb.GraphState()
    g.add_edges_from([(10, 30), (30, 20), (20, 40)])
    c = mb.MBQCircuit(g, input_nodes=[10], output_nodes=[40])
    assert list(c.graph.nodes()) == [0, 2, 1, 3]
    assert c.input_nodes == [0] and c.output_nodes == [3]
    assert c.measurement_order == [0, 2, 1, 3]
    assert c.trainable_nodes == [0, 2, 1]


def test_no_flow_and_errors():
    g = mb.GraphState([(0, 1), (1, 2), (2, 0)])
    c = mb.MBQCircuit(g, input_nodes=[0], output_nodes=[2])
    assert c.flow is None and c.measurement_order is None
    with pytest.raises(ValueError):
        mb.MBQCircuit(g, input_nodes=[0, 1], output_nodes=[2])
    with pytest.raises(KeyError):  # same as the reference: unknown label fails in the relabel map
        mb.MBQCircuit(g, input_nodes=[7], output_nodes=[2])
    with pytest.raises(ValueError):
        mb.MBQCircuit(g, input_nodes=[7], output_nodes=[2], relabel_indices=False)


def test_ment_contract():
    """mentpy tests/operators/test_ment.py:6-61 restated."""
    import numpy as np

    m = mb.Ment(0.5, "XY")
    assert (m.angle, m.plane) == (0.5, "XY") and not m.is_trainable()
    assert repr(m) == "Ment(0.5, XY)"
    assert mb.Ment("XY").is_trainable() and repr(mb.Ment("xy")) == "Ment(θ, XY)"
    assert mb.Ment("XY", 0.25).angle == 0.25
    assert not mb.Ment("X").is_trainable() and mb.Ment("Z").angle == 0
    assert np.allclose(mb.Ment(0.0).matrix(), np.array([[0, 1], [1, 0]]))
    assert np.allclose(mb.Ment(np.pi / 2, "XY").matrix(), np.array([[0, -1j], [1j, 0]]))
    with pytest.raises(ValueError):
        mb.Ment(plane="AB")
    with pytest.raises(ValueError):
        mb.Ment(0.3, "X")
    with pytest.raises(ValueError):
        mb.Ment("XY").matrix()
    with pytest.raises(ValueError):
        mb.Ment(0.3).matrix(0.4)
    with pytest.raises(TypeError):
        mb.Ment([0.1])
    for plane, mat in (("X", mb.gates.PauliX), ("Y", mb.gates.PauliY), ("Z", mb.gates.PauliZ)):
        assert np.allclose(mb.Ment(plane=plane).matrix(), mat)          # test_ment.py:58-61
    h = load_golden("helpers.json")
    from conftest import from_cplx

    for rec in h["ment"]:
        m = mb.Ment(rec["plane"]) if rec["angle"] is None else mb.Ment(rec["angle"], rec["plane"])
        assert np.array_equal(np.asarray(m.matrix(), dtype=complex), from_cplx(rec["matrix"]))
        p0, p1 = m.get_povm()
        assert np.array_equal(np.asarray(p0, dtype=complex), from_cplx(rec["p0"]))
        assert np.array_equal(np.asarray(p1, dtype=complex), from_cplx(rec["p1"]))
        assert mb.Ment(rec["plane"]).is_trainable() == rec["trainable"]


@pytest.mark.parametrize("n_layer", [1, 2])
@pytest.mark.parametrize("n_qubits", [4, 5])
@pytest.mark.parametrize("periodic", [True, False])
def test_spturb_trainable_nodes(n_layer, n_qubits, periodic):
    """The reference's own template test (tests/mbqc/test_mbqc_templates.py:12-22)."""
    spt = mb.templates.spturb(n_qubits, n_layer, periodic=periodic)
    blocks = n_qubits if periodic else n_qubits - 2
    assert len(spt.trainable_nodes) == n_qubits * n_layer + 2 * n_layer * blocks
    assert spt.flow is None and spt.measurement_order is None   # no causal flow: needs a user schedule
    with pytest.raises(ValueError):
        mb.templates.spturb(3, 1)


def test_controlled_measurements_host_layer():
    """ControlMent / MentOutcome (operators/controlled_ment.py:14-113, ment.py:13-120): the measurement
    order with the outcome dependencies, the trainable nodes and the tabulated conditions equal what
    the reference produced (tests/golden/dm_controlled.json); lowering refuses what the reference's
    simulator cannot run."""
    import warnings

    import mentpy_b200 as mb
    from mentpy_b200.plan import lower
    from oracle.gen_golden import CONTROL_CASES
    from oracle.pattern_data import PatternData

    for c in load_golden("dm_controlled.json")["cases"]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gs = CONTROL_CASES[c["name"]](mb, mb.ControlMent)
        assert gs.measurement_order == c["measurement_order"]
        assert gs.trainable_nodes == c["trainable_nodes"]
        import json
        assert json.loads(json.dumps(PatternData.from_circuit(gs).to_json())) == c["pattern"]
        plan = lower(gs, window_size=c["window_size"], mixed=True)
        for st in plan.steps:
            if st.node in gs.controlled_nodes:
                assert st.cond_mask and st.column == gs.trainable_nodes.index(st.node) and st.alt_angle_idx == st.column
        with pytest.raises(ValueError, match="only XY plane"):
            lower(gs, window_size=c["window_size"])  # state-vector path: like every non-XY node
    gs = mb.templates.linear_cluster(5)
    m = gs[0].outcome
    assert (~m)({0: 0}) and (m + 1)({0: 0}) and not (m * 1)({0: 0}) and ((m == 0) ^ m)({0: 1})
    gs[2] = mb.ControlMent(gs[0].outcome, 0.7, "XY", 0, "X")  # nothing trainable: the reference raises on run
    with pytest.raises(ValueError, match="not trainable"):
        lower(gs, mixed=True)
    gs[2] = mb.ControlMent(gs[0].outcome, 0.7, "XY", None, "XY")  # fixed TRUE branch: get_povm hands it the angle
    with pytest.raises(ValueError, match="fixed angle"):
        lower(gs, mixed=True)
    g2 = mb.templates.grid_cluster(2, 4)
    g2[2] = mb.ControlMent(g2[6].outcome == 0, None, "XY", 0, "X")  # reads a node measured later: the order adapts
    assert g2.measurement_order.index(6) < g2.measurement_order.index(2)
    with pytest.raises(AttributeError):
        gs[2] = mb.ControlMent(True, None, "XY", 0, "X")  # plain bool: mbqcircuit.py:617 fails the same way


def test_dev_mode_schedule_matches_reference():
    """dev_mode (np_simulator_sv.py:173-203, np_simulator_dm.py:160-201): the measurement order the
    window rule produces, for both simulators, against the reference (tests/golden/dev_mode.json)."""
    from mentpy_b200.plan import lower

    differs = 0
    for c in load_golden("dev_mode.json")["cases"]:
        name, args, kw = c["spec"]
        gs = getattr(mb.templates, name)(*args, **kw)
        for mixed, backend in ((False, "numpy-sv"), (True, "numpy-dm")):
            plan = lower(gs, mixed=mixed, window_size=c["window_size"], dev_mode=True, wires=c["wires"])
            assert plan.schedule_measure == c[backend]["order"]
            assert plan.window_nodes_after(0) == plan.schedule[: plan.window]
            assert sorted(plan.window_nodes_after(len(plan.steps))) == sorted(plan.output_nodes)
            differs += plan.schedule_measure != lower(gs, mixed=mixed, window_size=c["window_size"]).schedule_measure
    assert differs >= 4  # the fixture really exercises orders the plain schedule does not produce
    with pytest.raises(TypeError):
        lower(gs, dev_mode=True)
    with pytest.raises(ValueError, match="in no wire"):
        lower(gs, dev_mode=True, wires=[[0]])


def test_lowered_plans_of_the_new_fixtures_execute_to_the_reference_outputs():
    """The step records the kernels consume -- slots, masks, XYZ axes, condition masks / truth tables,
    dev_mode orders -- executed in numpy (tests/plan_emulator.py) against the outputs recorded from
    the reference: lowering errors surface on the CPU, before any kernel runs."""
    import warnings

    from conftest import from_cplx
    from mentpy_b200.plan import lower
    from oracle.gen_golden import CONTROL_CASES
    from plan_emulator import run_dm, run_sv

    def close(a, b, tol=1e-10):
        return np.abs(np.asarray(a) - np.asarray(b)).max() < tol

    for c in load_golden("dm_xyz_plane.json")["cases"]:
        name, args, kw = c["spec"]
        gs = getattr(mb.templates, name)(*args, **kw)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for v, a2 in c["xyz"].items():
                gs[int(v)] = mb.Ment(tuple(a2), "XYZ")
        n_in = len(gs.input_nodes)
        inp = np.full(2**n_in, 2.0 ** (-n_in / 2), dtype=complex) if c["input_state"] is None else from_cplx(c["input_state"])
        rho, oc = run_dm(lower(gs, window_size=c["window_size"], mixed=True), np.asarray(c["angles"]), inp)
        assert close(rho, from_cplx(c["output"]))
    for c in load_golden("dm_controlled.json")["cases"]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gs = CONTROL_CASES[c["name"]](mb, mb.ControlMent)
        n_in = len(gs.input_nodes)
        inp = np.full(2**n_in, 2.0 ** (-n_in / 2), dtype=complex) if c["input_state"] is None else from_cplx(c["input_state"])
        pl = lower(gs, window_size=c["window_size"], mixed=True)
        rho, oc = run_dm(pl, np.asarray(c["angles"]), inp)
        assert close(rho, from_cplx(c["output"])), c["name"]
        assert oc == [c["outcomes"][str(st.node)] for st in pl.steps]
    for c in load_golden("dev_mode.json")["cases"]:
        name, args, kw = c["spec"]
        gs = getattr(mb.templates, name)(*args, **kw)
        inp, ang = from_cplx(c["input_state"]), np.asarray(c["angles"])
        rho, _ = run_dm(lower(gs, window_size=c["window_size"], mixed=True, dev_mode=True, wires=c["wires"]), ang, inp)
        assert close(rho, from_cplx(c["numpy-dm"]["output"]))
        psi = run_sv(lower(gs, window_size=c["window_size"], dev_mode=True, wires=c["wires"]), ang, inp)
        want = from_cplx(c["numpy-sv"]["output"])  # run() default: |psi><psi|
        assert close(np.outer(psi, psi.conj()), want, 1e-9)
