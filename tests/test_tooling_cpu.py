"""CPU: pattern tooling (mentpy_b200/tooling) -- Pauli operators, GF(2) algebra, generalised flow, the
dynamical Lie algebra -- against the reference's own known answers (tests/operators/test_pauliop.py,
tests/utils/test_lie_algebra.py: restated here, the reference's versions need the `galois` package)
and against structural properties (gflow conditions, closure under commutators)."""
import itertools
import warnings

import numpy as np
import pytest

import mentpy_b200 as mb
from mentpy_b200 import tooling as tl
from mentpy_b200.mbqc.causal_flow import find_cflow


def test_pauli_op_known_answers():
    op = tl.PauliOp(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [1, 0, 1, 0]]))
    assert op.txt == "XI\nIX\nIZ\nYI"
    want = np.array([[1, 0, 0, 0, 0, 1], [0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 0, 1], [1, 1, 1, 1, 1, 1]])
    assert np.array_equal(tl.PauliOp("XIZ;ZII;IIZ;YYY").matrix, want)
    assert np.array_equal(tl.PauliOp(["XIZ", "ZII", "IIZ", "YYY"]).matrix, want)
    with pytest.raises(ValueError):
        tl.PauliOp(42)
    with pytest.raises(ValueError):
        tl.PauliOp("XI;XIZ")
    op = tl.PauliOp("XIZ;ZII;IIZ;IZI")
    assert [op[i].txt for i in range(4)] == ["XIZ", "ZII", "IIZ", "IZI"] and len(op) == 4
    assert op.get_subset([0, 2]).txt == "XZ\nZI\nIZ\nII"
    assert op[1:3].txt == "ZII\nIIZ" and op.number_of_qubits == 3
    a, b = tl.PauliOp("XI"), tl.PauliOp("ZI")
    assert (a * b).txt == "YI" and a.commutator(b).txt == "YI" and a.commutator(tl.PauliOp("IZ")) == 0
    assert a.symplectic_prod(b)[0, 0] == 1 and tl.PauliOp("XX").symplectic_prod(tl.PauliOp("ZZ"))[0, 0] == 0
    op.append(tl.PauliOp("YYY"))
    assert len(op) == 5 and tl.PauliOp("YYY") in op and tl.PauliOp("XXX") not in op
    assert tl.PauliOp("XZ") == tl.PauliOp(np.array([[1, 0, 0, 1]])) and hash(tl.PauliOp("XZ")) == hash(tl.PauliOp(["XZ"]))


def test_gf2_solver():
    rng = np.random.default_rng(0)
    for _ in range(50):
        m, n = rng.integers(1, 9, 2)
        a = rng.integers(0, 2, (m, n)).astype(np.uint8)
        x0 = rng.integers(0, 2, n).astype(np.uint8)
        b = (a.astype(int) @ x0) % 2
        x = tl.gf2_solve(a, b)
        assert x is not None and np.array_equal((a.astype(int) @ x) % 2, b)
        assert tl.gf2_rank(a) <= np.linalg.matrix_rank(a.astype(float))  # rank over GF(2) <= rank over R
    assert tl.gf2_solve(np.array([[1, 1], [1, 1]]), np.array([0, 1])) is None
    assert tl.gf2_rank(np.array([[1, 1, 0], [0, 1, 1], [1, 0, 1]])) == 2


@pytest.mark.parametrize("n_wires", [2, 3, 4])
def test_lie_algebra_grid(n_wires):
    """tests/utils/test_lie_algebra.py:5-12: grid clusters generate su(2^n) (+ the identity)."""
    gs = mb.templates.grid_cluster(n_wires, n_wires + 3)
    alg = tl.calculate_lie_algebra(gs, max_iter=10_000_000)
    assert len(alg) == tl.dim_su(2**n_wires) + 1
    # closed under commutators, no duplicates
    rows = {r.tobytes() for r in alg.matrix}
    assert len(rows) == len(alg)
    for i, j in itertools.islice(itertools.combinations(range(len(alg)), 2), 4000):
        c = alg[i].commutator(alg[j])
        assert c == 0 or c in alg


def test_lie_algebra_cylinder():
    """tests/utils/test_lie_algebra.py:15-24: the periodic grid with extra legs generates so(2^n)."""
    n = 4
    gs = mb.templates.grid_cluster(n, n + 3, periodic=True)
    legs = mb.templates.many_wires([2] * n)
    gs = mb.hstack((legs, gs, legs))
    alg = tl.calculate_lie_algebra(gs, max_iter=10_000_000)
    assert len(alg) == tl.dim_so(2**n) + 1
    assert tl.dim_sp(4) == 10 and tl.dim_su(4) == 15 and tl.dim_so(4) == 6


def test_generators_follow_the_stabilizer_constraints():
    gs = mb.templates.grid_cluster(2, 5)
    full = tl.calculate_complete_gens_lie_algebra(gs)
    n = len(gs.graph.nodes())
    index = {v: i for i, v in enumerate(gs.graph.nodes())}
    stabs = tl.graph_stabilizers(gs.graph)
    for row, j in zip(full.matrix, gs.outputc):
        # a product of stabilisers: commutes with every stabiliser
        assert not np.any(tl.PauliOp(row[None]).symplectic_prod(stabs))
        # Z on the measured nodes only at j; no X on nodes that are not after j
        assert [int(row[n + index[k]]) for k in gs.outputc] == [1 if k == j else 0 for k in gs.outputc]
        assert all(row[index[k]] == 0 for k in gs.measurement_order if k == j or not gs.partial_order(j, k))
    assert len(tl.calculate_gens_lie_algebra(gs)) <= len(gs.outputc)
    with pytest.raises(ValueError, match="Max iterations"):
        tl.lie_algebra_completion(tl.calculate_gens_lie_algebra(gs), max_iter=3)


def test_gflow_on_templates_and_beyond_causal_flow():
    for name, args in (("linear_cluster", [6]), ("grid_cluster", [2, 5]), ("grid_cluster", [3, 4]), ("many_wires", [[3, 4, 2]]),
                       ("muta", [2, 1])):
        gs = getattr(mb.templates, name)(*args)
        g, order, depth, layers = tl.find_gflow(gs.graph, gs.input_nodes, gs.output_nodes)
        assert tl.verify_gflow(gs.graph, gs.input_nodes, gs.output_nodes, g, layers)
        assert depth == max(layers.values()) and all(layers[v] == 0 for v in gs.output_nodes)
        for u in gs.outputc:
            assert all(order(u, v) for v in g(u))
    # graphs with a generalised flow but no causal flow (found by search: gflow strictly extends cflow)
    rng = np.random.default_rng(5)
    found = 0
    for _ in range(400):
        k = 3
        edges = [(i, k + j) for i in range(k) for j in range(k) if rng.random() < 0.6]
        if not edges:
            continue
        graph = mb.GraphState(edges)
        if len(graph.nodes()) != 2 * k:
            continue
        ins, outs = list(range(k)), list(range(k, 2 * k))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g, order, depth, layers = tl.find_gflow(graph, ins, outs)
            cf = find_cflow(graph, ins, outs)[0]
        if g is not None:
            assert tl.verify_gflow(graph, ins, outs, g, layers)
            if cf is None:
                found += 1
        else:
            assert cf is None  # no gflow => no causal flow
    assert found >= 1
    with pytest.warns(UserWarning, match="No gflow"):
        assert tl.find_gflow(mb.GraphState([(0, 2), (1, 2)]), [0, 1], [2])[0] is None


def test_haar_density_and_expressivity_from_samples():
    f = np.linspace(0, 1, 2001)
    for n in (1, 2, 3):
        dens = tl.haar_probability_density_of_fidelities(f, n)
        assert abs(np.trapezoid(dens, f) - 1) < 1e-3
    assert tl.haar_probability_density_of_fidelities(0.25, 2) == pytest.approx(3 * 0.75**2)
    gs = mb.templates.grid_cluster(2, 5)
    rng = np.random.default_rng(0)
    # fidelities of Haar-random 2-qubit states against a fixed one: F ~ Beta(1, N - 1)
    haar = rng.beta(1, 3, 200_000)
    peaked = np.clip(rng.normal(0.9, 0.02, 200_000), 0, 1)
    for method in ("KL", "RE", "JS"):
        e_haar = tl.expressivity_with_histogram(gs, n_bins=75, method=method, samples=haar)
        e_peak = tl.expressivity_with_histogram(gs, n_bins=75, method=method, samples=peaked)
        assert 0 <= e_haar < 0.01 < e_peak
    with pytest.raises(UserWarning):
        tl.expressivity_with_histogram(gs, method="nope", samples=haar)
