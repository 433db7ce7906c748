"""Numpy execution of a LoweredPlan exactly as the kernels index it (slots, recycled in place,
sign masks, input / output slot lists) -- lets the CPU suite check the host lowering against the
oracle without a GPU.  Straightforward loops; windows <= 6."""
import numpy as np

from mentpy_b200 import _lib


def _parity(x):
    return bin(x).count("1") & 1


def _seed(pl, inp):
    W, n_in = pl.window, len(pl.input_nodes)
    psi = np.zeros(2**W, complex)
    for i in range(2**W):
        src = 0
        for q, s in enumerate(pl.input_slot):
            src |= ((i >> s) & 1) << (n_in - 1 - q)
        sg = 0
        for a in range(W):
            if (i >> a) & 1:
                sg ^= _parity(i & pl.init_cz_mask[a])
        psi[i] = inp[src] * 2.0 ** (-(W - n_in) / 2) * (-1) ** sg
    return psi


def _cos_sin(st, ang):
    if st.angle_idx >= 0:
        return np.cos(ang[st.angle_idx]), np.sin(ang[st.angle_idx])
    return st.fixed_cos, st.fixed_sin


def run_sv(pl, ang, inp):
    """-> normalised output state [2^k] in pl.output_nodes order (global phase not tracked)."""
    W = pl.window
    psi = _seed(pl, inp)
    for st in pl.steps:
        c, s = _cos_sin(st, ang)
        e = c - 1j * s
        new = psi.copy()
        for i0 in range(2**W):
            if (i0 >> st.slot) & 1:
                continue
            t = psi[i0] + e * psi[i0 | (1 << st.slot)]
            new[i0] = t
            new[i0 | (1 << st.slot)] = t * ((-1) ** _parity(i0 & st.nbr_mask) if st.append else 1)
        psi = new / np.linalg.norm(new)
    k = len(pl.output_slot)
    out = np.zeros(2**k, complex)
    live = sum(1 << s for s in pl.output_slot)
    for i in range(2**W):
        if i & ~live:
            continue
        d = 0
        for q, s in enumerate(pl.output_slot):
            d |= ((i >> s) & 1) << (k - 1 - q)
        out[d] = psi[i]
    return out / np.linalg.norm(out)


def run_dm(pl, ang, inp):
    """-> (rho [2^k,2^k] in pl.output_nodes order, outcomes) with the reference's outcome rule."""
    W = pl.window
    psi = _seed(pl, inp)
    rho = np.outer(psi, psi.conj())
    outcomes = []
    for m, st in enumerate(pl.steps):
        c, s = _cos_sin(st, ang)
        plane, z = st.plane, st.fixed_z
        if st.cond_mask:  # outcome-controlled step: bit j of the mask = the outcome j + 1 measurements back
            bits = [j for j in range(32) if (st.cond_mask >> j) & 1]
            idx = sum(outcomes[m - 1 - j] << i for i, j in enumerate(bits))
            alt = (st.cond_table >> idx) & 1
            plane, z = (st.alt_plane, st.alt_z) if alt else (st.plane, st.fixed_z)
            aidx = st.alt_angle_idx if alt else st.angle_idx
            c, s = (np.cos(ang[aidx]), np.sin(ang[aidx])) if aidx >= 0 else ((st.alt_cos, st.alt_sin) if alt else (st.fixed_cos, st.fixed_sin))
        if plane == _lib.PLANE_XYZ:
            p00, p11, p10 = (1 + z) / 2, (1 - z) / 2, 0.5 * (c + 1j * s)
        elif plane == _lib.PLANE_XY:
            p00, p11, p10 = 0.5, 0.5, 0.5 * (c + 1j * s)
        elif plane == _lib.PLANE_XZ:
            p00, p11, p10 = (1 + s) / 2, (1 - s) / 2, 0.5 * c + 0j
        else:
            p00, p11, p10 = (1 + s) / 2, (1 - s) / 2, 0.5j * c
        bit = 1 << st.slot
        idx0 = [i for i in range(2**W) if not i & bit]
        sig0 = np.zeros((len(idx0), len(idx0)), complex)
        full = np.zeros_like(sig0)
        for a, r0 in enumerate(idx0):
            for b, c0 in enumerate(idx0):
                r00, r01, r10, r11 = rho[r0, c0], rho[r0, c0 | bit], rho[r0 | bit, c0], rho[r0 | bit, c0 | bit]
                sig0[a, b] = p00 * r00 + p11 * r11 + p10 * r01 + np.conj(p10) * r10
                full[a, b] = r00 + r11
        prob0 = np.real(np.trace(sig0)) / np.real(np.trace(full))
        outcome = 1 if prob0 < 1e-4 else 0
        outcomes.append(outcome)
        sig = full - sig0 if outcome else sig0
        sig = sig / np.real(np.trace(sig))
        new = np.zeros_like(rho)
        for a, r0 in enumerate(idx0):
            for b, c0 in enumerate(idx0):
                sr = (-1) ** _parity(r0 & st.nbr_mask) if st.append else 1
                sc = (-1) ** _parity(c0 & st.nbr_mask) if st.append else 1
                v = sig[a, b] / 2
                new[r0, c0], new[r0, c0 | bit] = v, v * sc
                new[r0 | bit, c0], new[r0 | bit, c0 | bit] = v * sr, v * sr * sc
        rho = new
    k = len(pl.output_slot)
    live = sum(1 << s for s in pl.output_slot)

    def widx(d):
        return sum(((d >> (k - 1 - q)) & 1) << s for q, s in enumerate(pl.output_slot))

    out = np.array([[rho[widx(r), widx(c)] for c in range(2**k)] for r in range(2**k)])
    return out / np.real(np.trace(out)), outcomes
