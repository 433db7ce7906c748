"""TEST INFRASTRUCTURE ONLY -- brute-force full-graph density-matrix simulation with noise.

Independent cross-check for the windowed noisy DM path (noise parity is otherwise UNPINNED: the
reference's only noise implementation is the PennyLane circuit of
mentpy/simulators/pennylane_simulator.py:97-166 and PennyLane is not installed).  This follows
that circuit's ORDER OF OPERATIONS literally, on all N qubits at once (so N <= ~11):

    prepare input on input_nodes, H|0> = |+> on the rest      (:113-119)
    CZ on every edge                                           (:120-121)
    the chosen single-qubit channel on EVERY wire              (:123-136)
    measure nodes in measurement order (outputs excluded)      (:138-153)

with the numpy simulators' outcome convention (always project on (I+M)/2, i.e. outcome 0 --
np_simulator_dm.py:335-338 -- so no byproduct corrections are needed) and returns the reduced
state of the output nodes in `quantum_output_nodes` order, big-endian.
"""
import numpy as np

from .matrix_free import kraus_ops, _projector
from .pattern_data import PatternData


def run_fullgraph_dm(pat: PatternData, angles, input_state=None, noise=None, noise_kwargs=None):
    N = pat.n_nodes
    if N > 12:
        raise ValueError("brute-force oracle limited to 12 nodes")
    angles = np.asarray(angles, dtype=float)
    n_in = len(pat.input_nodes)
    if input_state is None:
        input_state = np.full(2**n_in, 2.0 ** (-n_in / 2))
    # state vector over nodes 0..N-1 (axis v = node v)
    psi = np.asarray(input_state, dtype=complex).reshape([2] * n_in)
    others = [v for v in range(N) if v not in pat.input_nodes]
    for _ in others:
        psi = np.multiply.outer(psi, np.array([1.0, 1.0]) / np.sqrt(2))
    order_now = list(pat.input_nodes) + others
    psi = np.transpose(psi, [order_now.index(v) for v in range(N)])
    idx = np.indices([2] * N)
    for a, b in pat.edges:
        psi = psi * (1 - 2 * (idx[a] & idx[b]))
    rho = np.multiply.outer(psi, np.conj(psi))  # axes: rows 0..N-1, cols N..2N-1
    if noise:
        kr = kraus_ops(noise, **(noise_kwargs or {}))
        for v in range(N):
            acc = np.zeros_like(rho)
            for K in kr:
                u = np.moveaxis(np.tensordot(K, rho, axes=([1], [v])), 0, v)
                u = np.moveaxis(np.tensordot(np.conj(K), u, axes=([1], [N + v])), 0, N + v)
                acc += u
            rho = acc
    alive = list(range(N))
    measured = [v for v in pat.measurement_order if v not in pat.quantum_output_nodes]
    for node in measured:
        plane, fixed = pat.measurements[node]
        th = angles[pat.trainable_nodes.index(node)] if node in pat.trainable_nodes else fixed
        if plane == "X":
            plane, th = "XY", 0.0
        elif plane == "Y":
            plane, th = "XY", np.pi / 2
        p00, p11, p10 = (np.asarray(x).reshape(-1)[0] for x in _projector(plane, np.array([th])))
        P = np.array([[p00, np.conj(p10)], [p10, p11]])
        n = len(alive)
        pos = alive.index(node)
        # sigma = sum_ab P[b,a] rho_ab  == tr_node(P rho)
        t = np.tensordot(P, rho, axes=([1], [pos]))  # new axis 0 = row index of node
        t = np.moveaxis(t, 0, pos)
        t = np.trace(t, axis1=pos, axis2=n + pos)
        rho = t / np.real(np.trace(t.reshape(2 ** (n - 1), 2 ** (n - 1))))
        alive.remove(node)
    k = len(alive)
    perm = [alive.index(v) for v in pat.quantum_output_nodes]
    rho = np.transpose(rho, perm + [k + p for p in perm])
    return rho.reshape(2**k, 2**k)
