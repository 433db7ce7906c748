"""Dynamical Lie algebra of an MBQC pattern (the role of mentpy/utils/lie_algebra.py:18-185).

Every measured node j contributes one generator: the product of graph-state stabilisers that acts
trivially (no X part) on everything not measured after j, and whose Z parts on the measured nodes
reduce to Z_j alone; its restriction to the output qubits is the Pauli rotation axis that angle j
controls.  The algebra is the closure of those generators under commutators."""
from collections import deque
from typing import List

import numpy as np

from .gf2 import gf2_solve
from .pauli import PauliOp


def graph_stabilizers(graph) -> PauliOp:
    """K_v = X_v prod_{u ~ v} Z_u for every node, qubits in the graph's node order (graphstate.py:108-118)."""
    nodes = list(graph.nodes())
    index = {v: i for i, v in enumerate(nodes)}
    z = np.zeros((len(nodes), len(nodes)), dtype=np.uint8)
    for a, b in graph.edges():
        z[index[a], index[b]] = z[index[b], index[a]] = 1
    return PauliOp(np.hstack((np.eye(len(nodes), dtype=np.uint8), z)))


def _generator(j, state, stabs: PauliOp, index) -> PauliOp:
    n = len(index)
    rows, rhs = [], []
    for k in state.measurement_order:  # no X on nodes that are not strictly after j
        if k == j or not state.partial_order(j, k):
            rows.append(stabs.matrix[index[k], :n])
            rhs.append(0)
    for k in state.outputc:            # Z on the measured nodes: only on j itself
        rows.append(stabs.matrix[index[k], n:])
        rhs.append(1 if k == j else 0)
    x = gf2_solve(np.vstack(rows), np.asarray(rhs, dtype=np.uint8))
    if x is None:
        raise ValueError("Solution not found for j: " + str(j))
    op = np.zeros(2 * n, dtype=np.uint8)
    for i in np.nonzero(x)[0]:
        op ^= stabs.matrix[i]
    return PauliOp(op[None, :])


def calculate_complete_gens_lie_algebra(state) -> PauliOp:
    """One operator on ALL qubits per measured node, in the order of `state.outputc`."""
    stabs = graph_stabilizers(state.graph)
    index = {v: i for i, v in enumerate(state.graph.nodes())}
    ops = [_generator(j, state, stabs, index) for j in state.outputc]
    return PauliOp(np.vstack([o.matrix for o in ops]))


def remove_repeated_ops(ops: PauliOp) -> PauliOp:
    seen, keep = set(), []
    for r in ops.matrix:
        key = r.tobytes()
        if key not in seen:
            seen.add(key)
            keep.append(r)
    return PauliOp(np.vstack(keep))


def calculate_gens_lie_algebra(state) -> PauliOp:
    """The generators restricted to the output qubits, duplicates removed."""
    index = {v: i for i, v in enumerate(state.graph.nodes())}
    full = calculate_complete_gens_lie_algebra(state)
    return remove_repeated_ops(full.get_subset([index[v] for v in state.output_nodes]))


def lie_algebra_completion(generators: PauliOp, max_iter: int = 1000) -> PauliOp:
    """Close a set of Pauli operators under commutators (phases dropped); the identity is added at the
    end as in the reference (lie_algebra.py:147-150).  max_iter bounds the number of commutators tried."""
    rows: List[bytes] = []
    have = set()
    for r in generators.matrix:
        if r.tobytes() not in have:
            have.add(r.tobytes())
            rows.append(r.tobytes())
    width = generators.matrix.shape[1]
    n = width // 2

    def anticommute(a: np.ndarray, b: np.ndarray) -> bool:
        return bool((int(a[:n] @ b[n:]) + int(a[n:] @ b[:n])) & 1)

    mats = [np.frombuffer(r, dtype=np.uint8) for r in rows]
    queue = deque((i, j) for i in range(len(mats)) for j in range(i + 1, len(mats)))
    it = 0
    while queue:
        it += 1
        if it > max_iter:
            raise ValueError("Max iterations reached")
        i, j = queue.popleft()
        if not anticommute(mats[i], mats[j]):
            continue
        new = mats[i] ^ mats[j]
        if new.tobytes() in have:
            continue
        have.add(new.tobytes())
        mats.append(new)
        k = len(mats) - 1
        queue.extend((m, k) for m in range(k))
    ident = np.zeros(width, dtype=np.uint8)
    if ident.tobytes() not in have:
        mats.append(ident)
    return PauliOp(np.vstack(mats))


def calculate_lie_algebra(state, max_iter: int = 10000) -> PauliOp:
    return lie_algebra_completion(calculate_gens_lie_algebra(state), max_iter=max_iter)


def dim_su(n: int) -> int:
    return int(n**2 - 1)


def dim_so(n: int) -> int:
    return int(n * (n - 1) // 2)


def dim_sp(n: int) -> int:
    assert n % 2 == 0, "n must be even"
    h = n // 2
    return int(h * (2 * h + 1))
