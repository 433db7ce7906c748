"""TEST INFRASTRUCTURE ONLY -- matrix-free, batched numpy restatement of the reference hot path.

Checker only (tests/, smoke(), bench.py's cpu_baseline leg).  Same arithmetic as
oracle/dense_port.py / the reference, with every dense 2^n x 2^n operator replaced by its action:

  * projection onto (I+M)/2 of the window's first qubit, then the reference's pure-state
    "partial trace" (SUM over that qubit + renormalise)   np_simulator_sv.py:322-358,
    state_ops.py:42-74:
        red[r] = (1 + e^{i th})/2 * (psi[0,r] + e^{-i th} psi[1,r]);  red /= ||red||
  * kron with |+> (new qubit = LSB) and CZ with in-window neighbours   np_simulator_sv.py:207-223,
    gates.py:127-143:
        out[2r+b] = red[r]/sqrt2 * (-1)^{b * parity(r & nbr_mask)}
  * DM twin   np_simulator_dm.py:151-216, :307-346, state_ops.py:77-119:
        sigma = sum_ab P[b,a] rho_ab / prob,  prob = Re tr(rho P),  outcome 1 iff prob0 < 1e-4
  * single-qubit noise channels placed where mentpy/simulators/pennylane_simulator.py:118-136 puts
    them (after every CZ touching the qubit, before its measurement; on the output qubits at the
    end).  Kraus sets are PennyLane's published definitions (pennylane >= 0.30, unpinned in the
    reference's requirements.txt:6, not installed here): PARITY UNPINNED for noise -- pinned only
    by oracle/fullgraph_noise.py (independent brute force) and invariants.

Layout differs on purpose from the CUDA product (which recycles bit slots in place): here the
window is physically shifted like the reference does, so agreement is not by construction.
All arrays carry a leading batch axis B.
"""
import numpy as np

from .pattern_data import PatternData

SQRT1_2 = 1.0 / np.sqrt(2.0)


def _parity(x):
    x = np.asarray(x, dtype=np.uint64)
    for s in (32, 16, 8, 4, 2, 1):
        x = x ^ (x >> np.uint64(s))
    return (x & np.uint64(1)).astype(np.int64)


def _plan(pat: PatternData, window_size, schedule, mixed):
    out_excl = pat.quantum_output_nodes if mixed else pat.output_nodes
    if schedule is None:
        schedule = list(pat.measurement_order)
        if window_size == 1:
            window_size = len(pat.input_nodes) + 1
    sched_meas = [v for v in schedule if v not in out_excl]
    return list(schedule), sched_meas, window_size


def _perm_axes(state_axes_nodes, target_nodes):
    return [state_axes_nodes.index(v) for v in target_nodes]


def _seed(pat, schedule, w, input_states, B):
    """input (in input_nodes order) -> schedule[:|I|] order, x |+>^(w-|I|), initial CZ signs.
    np_simulator_sv.py:81-128."""
    n_in = len(pat.input_nodes)
    if input_states is None:
        st = np.full((1, 2**n_in), 2.0 ** (-n_in / 2), dtype=complex)
    else:
        st = np.asarray(input_states, dtype=complex)
        if st.ndim == 1:
            st = st[None, :]
    nb = st.shape[0]
    if n_in > 0:
        t = st.reshape([nb] + [2] * n_in)
        axes = [0] + [1 + pat.input_nodes.index(v) for v in schedule[:n_in]]
        st = t.transpose(axes).reshape(nb, -1)
    for _ in range(w - n_in):
        st = np.kron(st, np.array([[SQRT1_2, SQRT1_2]]))
    first = schedule[:w]
    idx = np.arange(2**w, dtype=np.uint64)
    sign = np.zeros(2**w, dtype=np.int64)
    for a, b in pat.edges:
        if a in first and b in first:
            ba = (idx >> np.uint64(w - 1 - first.index(a))) & np.uint64(1)
            bb = (idx >> np.uint64(w - 1 - first.index(b))) & np.uint64(1)
            sign ^= (ba & bb).astype(np.int64)
    st = st * (1 - 2 * sign)[None, :]
    if nb == 1 and B > 1:
        st = np.repeat(st, B, axis=0)
    return st


def _angles_for_step(pat, node, angles):
    plane, fixed = pat.measurements[node]
    if node in pat.trainable_nodes:
        return plane, angles[:, pat.trainable_nodes.index(node)]
    if plane == "X":
        return "XY", np.zeros(angles.shape[0])
    if plane == "Y":
        return "XY", np.full(angles.shape[0], np.pi / 2)
    if plane == "Z":
        return "Z", np.zeros(angles.shape[0])
    if plane == "XYZ":  # two fixed angles (ment.py:239-251): [B,2]
        return plane, np.tile(np.asarray(fixed, dtype=float), (angles.shape[0], 1))
    return plane, np.full(angles.shape[0], fixed)


def _nbr_mask(pat, new_node, window_nodes):
    """bit mask over the reduced index r (window positions 0..n-2 -> bits n-2..0)."""
    n = len(window_nodes)
    mask = 0
    for nb in pat.neighbors(new_node):
        if nb in window_nodes[:-1]:
            mask |= 1 << (n - 2 - window_nodes.index(nb))
    return mask


def run_sv_batch(pat, angles, input_states=None, window_size=1, schedule=None, output_form="sv"):
    """Batched restatement of NumpySimulatorSV.run (np_simulator_sv.py:227-297).
    angles [B,T] -> [B,2^k] ('sv') or [B,2^k,2^k] ('dm')."""
    angles = np.atleast_2d(np.asarray(angles, dtype=float))
    B = angles.shape[0]
    schedule, sched_meas, w = _plan(pat, window_size, schedule, mixed=False)
    psi = _seed(pat, schedule, w, input_states, B)
    N = pat.n_nodes
    for cm0, node in enumerate(sched_meas):
        _, th = _angles_for_step(pat, node, angles)
        half = psi.shape[1] // 2
        e = np.exp(-1j * th)[:, None]
        red = (1 + np.conj(e)) / 2 * (psi[:, :half] + e * psi[:, half:])
        red = red / np.linalg.norm(red, axis=1)[:, None]
        cm = cm0 + 1
        if cm + w <= N:
            win = schedule[cm : cm + w]
            mask = _nbr_mask(pat, win[-1], win)
            sgn = 1 - 2 * _parity(np.arange(half, dtype=np.uint64) & np.uint64(mask))
            out = np.empty((B, 2 * half), dtype=complex)
            out[:, 0::2] = red * SQRT1_2
            out[:, 1::2] = red * SQRT1_2 * sgn[None, :]
            psi = out
        else:
            psi = red
    current = schedule[len(sched_meas) : len(sched_meas) + w]
    k = len(current)
    if pat.quantum_output_nodes != current:
        t = psi.reshape([B] + [2] * k)
        axes = [0] + [1 + current.index(v) for v in pat.output_nodes]
        psi = t.transpose(axes).reshape(B, -1)
    if output_form == "dm":
        return psi[:, :, None] * np.conj(psi[:, None, :])
    return psi


# ---------------------------------------------------------------------------------------------
# noise channels: rho -> sum_k K rho K^+ on one qubit, in block form on (rho00, rho01, rho10, rho11)
# ---------------------------------------------------------------------------------------------
def kraus_ops(kind, p=0.0, gamma=None, p_gad=None):
    """PennyLane's published Kraus sets for the channels named at pennylane_simulator.py:125-134."""
    if kind is None or kind == "none":
        return [np.eye(2, dtype=complex)]
    x = np.array([[0, 1], [1, 0]], dtype=complex)
    y = np.array([[0, -1j], [1j, 0]], dtype=complex)
    z = np.array([[1, 0], [0, -1]], dtype=complex)
    g = p if gamma is None else gamma
    if kind == "depolarizing":
        return [np.sqrt(1 - p) * np.eye(2), np.sqrt(p / 3) * x, np.sqrt(p / 3) * y, np.sqrt(p / 3) * z]
    if kind == "phase_flip":
        return [np.sqrt(1 - p) * np.eye(2), np.sqrt(p) * z]
    if kind == "bit_flip":
        return [np.sqrt(1 - p) * np.eye(2), np.sqrt(p) * x]
    if kind == "amplitude_damping":
        return [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex),
                np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex)]
    if kind == "phase_damping":
        return [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex),
                np.array([[0, 0], [0, np.sqrt(g)]], dtype=complex)]
    if kind == "generalized_amplitude_damping":
        q = 0.5 if p_gad is None else p_gad
        return [np.sqrt(q) * np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex),
                np.sqrt(q) * np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex),
                np.sqrt(1 - q) * np.array([[np.sqrt(1 - g), 0], [0, 1]], dtype=complex),
                np.sqrt(1 - q) * np.array([[0, 0], [np.sqrt(g), 0]], dtype=complex)]
    raise ValueError(f"Unrecognized circuit noise: {kind}")


def apply_channel_first(rho, kraus):
    """Channel on window position 0 (MSB) of rho [B,D,D]."""
    B, D, _ = rho.shape
    h = D // 2
    blk = rho.reshape(B, 2, h, 2, h)
    out = np.zeros_like(blk)
    for K in kraus:
        out += np.einsum("ia,bajcl,kc->bijkl", K, blk, np.conj(K))
    return out.reshape(B, D, D)


def apply_channel_at(rho, kraus, pos, n):
    """Channel on window position `pos` of an n-qubit rho [B,2^n,2^n]."""
    B = rho.shape[0]
    t = rho.reshape([B] + [2] * (2 * n))
    out = np.zeros_like(t)
    for K in kraus:
        u = np.tensordot(K, t, axes=([1], [1 + pos]))
        u = np.moveaxis(u, 0, 1 + pos)
        u = np.tensordot(np.conj(K), u, axes=([1], [1 + n + pos]))
        u = np.moveaxis(u, 0, 1 + n + pos)
        out += u
    return out.reshape(B, 2**n, 2**n)


def _projector(plane, th):
    """(I+M)/2 per batch element -> p00[B], p11[B], p10[B] (p01 = conj p10).  ment.py:228-260."""
    if plane == "XYZ":  # M = cos t1 cos t2 X + sin t1 cos t2 Y + sin t2 Z
        t1, t2 = th[:, 0], th[:, 1]
        nz = np.sin(t2)
        return (1 + nz) / 2, (1 - nz) / 2, 0.5 * np.cos(t2) * (np.cos(t1) + 1j * np.sin(t1))
    c, s = np.cos(th), np.sin(th)
    if plane == "XY":
        return np.full_like(c, 0.5), np.full_like(c, 0.5), 0.5 * (c + 1j * s)
    if plane == "XZ":
        return (1 + s) / 2, (1 - s) / 2, 0.5 * c + 0j
    if plane == "YZ":
        return (1 + s) / 2, (1 - s) / 2, 0.5j * c
    if plane == "Z":
        return np.ones_like(c), np.zeros_like(c), np.zeros_like(c) + 0j
    raise NotImplementedError(f"plane {plane}")


def run_dm_batch(pat, angles, input_states=None, window_size=1, schedule=None,
                 noise=None, noise_kwargs=None, return_outcomes=False, mode="sample", z_outcomes=None):
    """Batched restatement of NumpySimulatorDM.run (np_simulator_dm.py:218-283) + optional noise.
    angles [B,T] -> rho [B,2^k,2^k].  Plane-Z nodes: mode="expectation" traces the qubit out
    unprojected and records prob1 as the (float) outcome (np_simulator_dm.py:327-344); in
    mode="sample" the reference draws them at random even under force0 (np_simulator_dm.py:329-346):
    pass the drawn record as z_outcomes [B, n_measurements] (entries of the plane-Z steps are used) to
    get that branch -- projection on |0><0| / |1><1|; `outcomes` then also returns prob1 of those
    steps in a second array."""
    angles = np.atleast_2d(np.asarray(angles, dtype=float))
    B = angles.shape[0]
    schedule, sched_meas, w = _plan(pat, window_size, schedule, mixed=True)
    psi = _seed(pat, schedule, w, input_states, B)
    rho = psi[:, :, None] * np.conj(psi[:, None, :])
    kr = kraus_ops(noise, **(noise_kwargs or {})) if noise else None
    N = pat.n_nodes
    has_z = any(pat.measurements[v][0] == "Z" for v in sched_meas)
    outcomes = np.zeros((B, len(sched_meas)), dtype=np.float64 if has_z else np.int8)
    zsample = has_z and mode not in ("expectation", "exp")
    if zsample and z_outcomes is None:
        raise NotImplementedError("plane Z in mode='sample' is random in the reference: pass z_outcomes")
    zprob1 = np.zeros((B, len(sched_meas)))
    for cm0, node in enumerate(sched_meas):
        plane, th = _angles_for_step(pat, node, angles)
        if kr is not None:
            rho = apply_channel_first(rho, kr)
        D = rho.shape[1]
        h = D // 2
        if node in pat.controls:
            # controlled_ment.py:96-113: per row, the condition on the outcomes so far picks the branch
            ctl = pat.controls[node]
            col = angles[:, pat.trainable_nodes.index(node)]
            idx = np.zeros(B, dtype=int)
            for i, r in enumerate(ctl["reads"]):
                idx |= (outcomes[:, sched_meas.index(r)].astype(int) & 1) << i
            take = np.asarray(ctl["table"], dtype=bool)[idx]
            parts = []
            for pl, fixed in (ctl["false"], ctl["true"]):
                if pl in ("X", "Y"):
                    pl, fixed = "XY", (0.0 if pl == "X" else np.pi / 2)
                ang = col if fixed is None else (np.tile(np.asarray(fixed, dtype=float), (B, 1)) if pl == "XYZ" else np.full(B, fixed))
                parts.append(_projector(pl, ang))
            p00, p11, p10 = (np.where(take, t, f) for f, t in zip(*parts))
        else:
            p00, p11, p10 = _projector(plane, th)
        r00, r01, r10, r11 = rho[:, :h, :h], rho[:, :h, h:], rho[:, h:, :h], rho[:, h:, h:]
        sig0 = (p00[:, None, None] * r00 + p11[:, None, None] * r11
                + p10[:, None, None] * r01 + np.conj(p10)[:, None, None] * r10)
        full = r00 + r11
        prob0 = np.real(np.trace(sig0, axis1=1, axis2=2))
        prob1 = np.real(np.trace(full, axis1=1, axis2=2)) - prob0
        if plane == "Z" and zsample:  # P0 = |0><0| (with the channel in front), outcome given
            sig0 = r00  # the channel has already acted on rho (apply_channel_first above)
            prob0 = np.real(np.trace(sig0, axis1=1, axis2=2))
            prob1 = np.real(np.trace(full, axis1=1, axis2=2)) - prob0
            take1 = np.asarray(z_outcomes)[:, cm0].astype(bool)
            outcomes[:, cm0] = take1
            zprob1[:, cm0] = prob1 / (prob0 + prob1)
            sig = np.where(take1[:, None, None], (full - sig0) / np.where(take1, prob1, 1.0)[:, None, None],
                           sig0 / np.where(take1, 1.0, prob0)[:, None, None])
        elif plane == "Z":  # expectation mode: no projection, outcome = prob1 / (prob0 + prob1)
            outcomes[:, cm0] = prob1 / (prob0 + prob1)
            sig = full
        else:
            take1 = prob0 < 1e-4
            outcomes[:, cm0] = take1
            sig = np.where(take1[:, None, None], (full - sig0) / np.where(take1, prob1, 1.0)[:, None, None],
                           sig0 / np.where(take1, 1.0, prob0)[:, None, None])
        cm = cm0 + 1
        if cm + w <= N:
            win = schedule[cm : cm + w]
            mask = _nbr_mask(pat, win[-1], win)
            sgn = 1 - 2 * _parity(np.arange(h, dtype=np.uint64) & np.uint64(mask))
            v = np.stack([np.ones(h), sgn.astype(float)], axis=1).reshape(-1)  # [2h]: (+1, sgn_r)
            big = np.repeat(np.repeat(sig, 2, axis=1), 2, axis=2) * 0.5
            rho = big * v[None, :, None] * v[None, None, :]
        else:
            rho = sig
    current = [v for v in schedule if v not in sched_meas]
    k = len(current)
    if kr is not None:
        for pos in range(k):
            rho = apply_channel_at(rho, kr, pos, k)
    if pat.quantum_output_nodes != current:
        t = rho.reshape([B] + [2] * (2 * k))
        perm = [current.index(v) for v in pat.quantum_output_nodes]
        axes = [0] + [1 + p for p in perm] + [1 + k + p for p in perm]
        rho = t.transpose(axes).reshape(B, 2**k, 2**k)
    if return_outcomes:
        return (rho, outcomes, zprob1) if zsample else (rho, outcomes)
    return rho


def linear_cluster_analytic(angles, input_state=None):
    """Closed form for linear_cluster(L): out ~ J(-th_{L-2}) ... J(-th_0)|in>,
    J(a) = [[1, e^{ia}], [1, -e^{ia}]]/sqrt2, any window size."""
    angles = np.atleast_2d(np.asarray(angles, dtype=float))
    B = angles.shape[0]
    v = np.tile(np.array([SQRT1_2, SQRT1_2], dtype=complex) if input_state is None
                else np.asarray(input_state, dtype=complex), (B, 1))
    for j in range(angles.shape[1]):
        e = np.exp(-1j * angles[:, j])
        a, b = v[:, 0], v[:, 1]
        v = np.stack([(a + e * b) * SQRT1_2, (a - e * b) * SQRT1_2], axis=1)
    return v / np.linalg.norm(v, axis=1)[:, None]
