"""TEST INFRASTRUCTURE ONLY -- outcome sampling with flow corrections ("force0=False"), CPU side.

The numpy simulators of the reference raise NotImplementedError for force0=False
(np_simulator_sv.py:50-51) and Flow.adapt_angle(s) are stubs (mbqc/flow.py:105-109); the only
place the reference spells out what a non-zero outcome does is its PennyLane circuit
(mentpy/simulators/pennylane_simulator.py:138-153):

    for node in measurement_order (outputs excluded):
        RZ(-angle); H; m = measure(node)
        if m: X on f(node); Z on every neighbour n of f(node), n != node, measured after node

PennyLane is not installed (parity UNPINNED), so two independent restatements are kept here and
checked against each other and against the CUDA kernels:

  * run_fullgraph_branch / branch_average: brute force on all N qubits, the corrections applied
    PHYSICALLY as Pauli gates in exactly that order (N <= 11).  branch_average is what the
    PennyLane backend returns (deferred measurements: outcome-averaged density matrix).
  * feedforward_tables + run_sv_sampled: the windowed formulation the kernels use -- a pending
    X^a Z^b on a qubit measured in the XY plane is folded into its angle,
    theta' = (-1)^a theta + b pi, and into a final X^a Z^b on the output qubits.

philox4x32_10 restates the published Random123 generator (Salmon et al., SC'11; constants
0xD2511F53, 0xCD9E8D57, Weyl 0x9E3779B9, 0xBB67AE85) and is pinned by its known-answer vectors
in tests/test_feedforward_cpu.py.
"""
import itertools

import numpy as np

from .matrix_free import SQRT1_2, _nbr_mask, _parity, _plan, _seed, kraus_ops
from .pattern_data import PatternData

# ---- Philox4x32-10 ------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [np.asarray(counter)[..., i].astype(np.uint64) for i in range(4)]
    k0 = np.asarray(key)[..., 0].astype(np.uint64)
    k1 = np.asarray(key)[..., 1].astype(np.uint64)
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & _LO, p1 & _LO, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & _LO, p0 & _LO]
        k0 = (k0 + np.uint64(_W0)) & _LO
        k1 = (k1 + np.uint64(_W1)) & _LO
    return np.stack(c, axis=-1).astype(np.uint32)


def uniform_for(seed: int, sample, step: int):
    """The uniform in [0, 1) the kernels draw for (sample, measurement step): counter =
    (sample lo, sample hi, step, 0), key = (seed lo, seed hi); 53 bits from words 0 and 1."""
    sample = np.asarray(sample, dtype=np.uint64)
    ctr = np.stack([sample & _LO, sample >> np.uint64(32), np.full_like(sample, step), np.zeros_like(sample)], axis=-1)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64), ctr.shape[:-1] + (2,))
    x = philox4x32_10(ctr, key).astype(np.uint64)
    return ((x[..., 0] >> np.uint64(5)) * np.uint64(1 << 26) + (x[..., 1] >> np.uint64(6))).astype(np.float64) * 2.0 ** -53


# ---- brute force, corrections applied as gates ---------------------------------------------------
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)


def _apply_1q(rho, U, v, N):
    rho = np.moveaxis(np.tensordot(U, rho, axes=([1], [v])), 0, v)
    return np.moveaxis(np.tensordot(np.conj(U), rho, axes=([1], [N + v])), 0, N + v)


def _node_angle(pat, node, angles):
    plane, fixed = pat.measurements[node]
    if plane == "X":
        return 0.0
    if plane == "Y":
        return np.pi / 2
    if plane != "XY":
        raise ValueError("flow corrections are defined for XY-plane measurements")
    return float(angles[pat.trainable_nodes.index(node)]) if node in pat.trainable_nodes else float(fixed)


def _full_rho(pat, input_state, noise, noise_kwargs):
    N = pat.n_nodes
    if N > 11:
        raise ValueError("brute-force oracle limited to 11 nodes")
    n_in = len(pat.input_nodes)
    if input_state is None:
        input_state = np.full(2**n_in, 2.0 ** (-n_in / 2))
    psi = np.asarray(input_state, dtype=complex).reshape([2] * n_in)
    others = [v for v in range(N) if v not in pat.input_nodes]
    for _ in others:
        psi = np.multiply.outer(psi, np.array([1.0, 1.0]) / np.sqrt(2))
    order_now = list(pat.input_nodes) + others
    psi = np.transpose(psi, [order_now.index(v) for v in range(N)])
    idx = np.indices([2] * N)
    for a, b in pat.edges:
        psi = psi * (1 - 2 * (idx[a] & idx[b]))
    rho = np.multiply.outer(psi, np.conj(psi))
    if noise:
        kr = kraus_ops(noise, **(noise_kwargs or {}))
        for v in range(N):
            rho = sum(_apply_1q(rho, K, v, N) for K in kr)
    return rho


def run_fullgraph_branch(pat: PatternData, flow, angles, outcomes, input_state=None, noise=None, noise_kwargs=None):
    """One outcome record, pennylane_simulator.py:138-153 literally.  Returns (branch probability,
    normalised output density matrix in quantum_output_nodes order, big-endian)."""
    N = pat.n_nodes
    rho = _full_rho(pat, input_state, noise, noise_kwargs)
    order = pat.measurement_order
    measured = [v for v in order if v not in pat.quantum_output_nodes]
    prob = 1.0
    for node, m in zip(measured, outcomes):
        th = _node_angle(pat, node, angles)
        bra = np.array([1.0, (-1) ** int(m) * np.exp(-1j * th)]) / np.sqrt(2)  # <m| H RZ(-th), up to a phase
        P = np.outer(np.conj(bra), bra)  # projector |b><b|, the qubit is left in |b>
        rho = _apply_1q(rho, P, node, N)
        p = float(np.real(np.trace(rho.reshape(2**N, 2**N))))
        prob *= p
        if p <= 0:
            return 0.0, None
        rho = rho / p
        if m:
            tgt = flow[node]
            rho = _apply_1q(rho, _X, tgt, N)
            for nb in pat.neighbors(tgt):
                if nb != node and order.index(nb) > order.index(node):
                    rho = _apply_1q(rho, _Z, nb, N)
    keep = list(pat.quantum_output_nodes)
    mat = rho
    axes_alive = list(range(N))
    for v in measured:  # trace out the measured qubits
        pos = axes_alive.index(v)
        n = len(axes_alive)
        mat = np.trace(mat, axis1=pos, axis2=n + pos)
        axes_alive.remove(v)
    k = len(axes_alive)
    perm = [axes_alive.index(v) for v in keep]
    mat = np.transpose(mat, perm + [k + p for p in perm]).reshape(2**k, 2**k)
    return prob, mat


def branch_average(pat, flow, angles, input_state=None, noise=None, noise_kwargs=None):
    """Outcome-averaged output state = what the reference's PennyLane backend computes."""
    n_meas = len([v for v in pat.measurement_order if v not in pat.quantum_output_nodes])
    acc, total = 0.0, 0.0
    for rec in itertools.product((0, 1), repeat=n_meas):
        p, rho = run_fullgraph_branch(pat, flow, angles, rec, input_state, noise, noise_kwargs)
        if p > 0:
            acc = acc + p * rho
            total += p
    return acc, total


# ---- windowed formulation (what the kernels do) ---------------------------------------------------
def feedforward_tables(pat: PatternData, flow):
    """For every measured node j: (x_sources, z_sources) = the earlier measured nodes whose outcome
    toggles a pending X / Z on j; same for the output nodes.  Direct transcription of the
    correction rule above."""
    order = pat.measurement_order
    pos = {v: i for i, v in enumerate(order)}
    xs = {v: [] for v in range(pat.n_nodes)}
    zs = {v: [] for v in range(pat.n_nodes)}
    for node in order:
        if node in pat.quantum_output_nodes:
            continue
        tgt = flow[node]
        xs[tgt].append(node)
        for nb in pat.neighbors(tgt):
            if nb != node and pos[nb] > pos[node]:
                zs[nb].append(node)
    return xs, zs


def run_sv_sampled(pat: PatternData, flow, angles, seed=0, sample_offset=0, input_states=None, window_size=1,
                   schedule=None, forced=None, correct=True):
    """Batched windowed run with sampled (Philox, see uniform_for) or forced outcomes and adapted
    angles.  Returns (psi_out [B,2^k] normalised, outcomes [B,M] int8, branch probability [B],
    byproducts (x_bits, z_bits) [B,k] for the output nodes in quantum_output_nodes order)."""
    angles = np.atleast_2d(np.asarray(angles, dtype=float))
    B = angles.shape[0]
    schedule, sched_meas, w = _plan(pat, window_size, schedule, mixed=False)
    psi = _seed(pat, schedule, w, input_states, B)
    xs, zs = feedforward_tables(pat, flow)
    N = pat.n_nodes
    M = len(sched_meas)
    outcomes = np.zeros((B, M), dtype=np.int8)
    rec = {}
    prob = np.ones(B)
    for cm0, node in enumerate(sched_meas):
        base = np.array([_node_angle(pat, node, angles[b]) for b in range(B)])
        a = sum((rec[i] for i in xs[node]), np.zeros(B, dtype=np.int64)) % 2
        bz = sum((rec[i] for i in zs[node]), np.zeros(B, dtype=np.int64)) % 2
        th = (1 - 2 * a) * base + bz * np.pi
        half = psi.shape[1] // 2
        e = np.exp(-1j * th)[:, None]
        t0 = psi[:, :half] + e * psi[:, half:]
        t1 = psi[:, :half] - e * psi[:, half:]
        n0 = np.sum(np.abs(t0) ** 2, axis=1)
        n1 = np.sum(np.abs(t1) ** 2, axis=1)
        p0 = n0 / (n0 + n1)
        if forced is None:
            m = (~(uniform_for(seed, np.arange(B, dtype=np.uint64) + np.uint64(sample_offset), cm0) < p0)).astype(np.int64)
        else:
            m = np.asarray(forced)[:, cm0].astype(np.int64)
        prob = prob * np.where(m == 1, 1 - p0, p0)
        red = np.where(m[:, None] == 1, t1, t0)
        nr = np.sqrt(np.sum(np.abs(red) ** 2, axis=1))[:, None]
        with np.errstate(invalid="ignore", divide="ignore"):
            red = red / nr
        outcomes[:, cm0] = m
        rec[node] = m
        cm = cm0 + 1
        if cm + w <= N:  # append |+> and CZ with the in-window neighbours (np_simulator_sv.py:207-223)
            win = schedule[cm: cm + w]
            mask = _nbr_mask(pat, win[-1], win)
            sgn = 1 - 2 * _parity(np.arange(half, dtype=np.uint64) & np.uint64(mask))
            nxt = np.empty((B, 2 * half), dtype=complex)
            nxt[:, 0::2] = red * SQRT1_2
            nxt[:, 1::2] = red * SQRT1_2 * sgn[None, :]
            psi = nxt
        else:
            psi = red
    live = schedule[len(sched_meas):]
    k = len(live)
    perm = [live.index(v) for v in pat.quantum_output_nodes]
    out = psi.reshape([B] + [2] * k).transpose([0] + [1 + p for p in perm]).reshape(B, -1)
    xb = np.stack([sum((rec[i] for i in xs[v]), np.zeros(B, dtype=np.int64)) % 2 for v in pat.quantum_output_nodes], axis=1)
    zb = np.stack([sum((rec[i] for i in zs[v]), np.zeros(B, dtype=np.int64)) % 2 for v in pat.quantum_output_nodes], axis=1)
    if correct:  # Z^zb then X^xb on every output qubit (first node = MSB)
        idx = np.arange(2**k)
        for q in range(k):
            bit = k - 1 - q
            sgn = 1 - 2 * ((idx >> bit) & 1)[None, :] * zb[:, q][:, None]
            out = out * sgn
            flipped = out[:, idx ^ (1 << bit)]
            out = np.where(xb[:, q][:, None] == 1, flipped, out)
    return out, outcomes, prob, (xb, zb)
