"""State helpers with the reference's calculator API (mentpy/calculator/state_ops.py:16-156),
evaluated by CUDA kernels (csrc/calc.cuh).  numpy in -> numpy out, CUDA tensor in -> CUDA tensor
out.  Note the reference's pure-state "partial trace" is a SUM over the traced qubits followed by
a renormalisation (state_ops.py:66-73), not a projection -- reproduced as is."""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ["partial_trace", "pure2density", "partial_trace_pure_state", "partial_trace_mixed_state", "fidelity"]


def _to_device(data):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("mentpy_b200 needs a CUDA device: there is no CPU fallback.")
    if isinstance(data, torch.Tensor):
        return data.to(device="cuda", dtype=torch.complex128).contiguous(), False
    return torch.from_numpy(np.ascontiguousarray(data, dtype=np.complex128)).cuda(), True


def _finish(t, on_host):
    return t.cpu().numpy() if on_host else t


def _n_qubits(dim):
    n = int(round(np.log2(dim)))
    if 2**n != dim:
        raise ValueError("Invalid input shape for quantum state.")
    return n


def pure2density(psi):
    import torch

    d, on_host = _to_device(psi)
    n = _n_qubits(d.shape[0])
    out = torch.empty((2**n, 2**n), dtype=torch.complex128, device=d.device)
    _lib.check(_lib.load().mbqc_pure2density(d.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return _finish(out, on_host)


def partial_trace_pure_state(psi, indices):
    import torch

    d, on_host = _to_device(psi)
    n = _n_qubits(d.shape[0])
    idx = (C.c_int32 * max(len(indices), 1))(*[int(i) for i in indices])
    out = torch.empty(2 ** (n - len(set(indices))), dtype=torch.complex128, device=d.device)
    _lib.check(_lib.load().mbqc_partial_trace_pure(d.data_ptr(), n, idx, len(indices), out.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream))
    return _finish(out, on_host)


def partial_trace_mixed_state(rho, indices):
    import torch

    d, on_host = _to_device(rho)
    n = _n_qubits(d.shape[0])
    idx = (C.c_int32 * max(len(indices), 1))(*[int(i) for i in indices])
    k = n - len(set(indices))
    out = torch.empty((2**k, 2**k), dtype=torch.complex128, device=d.device)
    _lib.check(_lib.load().mbqc_partial_trace_mixed(d.data_ptr(), n, idx, len(indices), out.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream))
    return _finish(out, on_host)


def partial_trace(data, indices):
    shape = tuple(data.shape)
    if len(shape) == 1:
        return partial_trace_pure_state(data, indices)
    if len(shape) == 2:
        return partial_trace_mixed_state(data, indices)
    raise ValueError("Invalid input shape for quantum state.")


def fidelity(a, b):
    """Uhlmann fidelity (tr sqrt(sqrt(a) b sqrt(a)))^2 of two density matrices -- what
    mentpy.calculator.fidelity (calculator/borrows.py:4-6, PennyLane's math.fidelity) returns;
    the loss of docs/tutorials/intro-to-mbqml.rst:35-42 calls it with a pure target.  State
    vectors are accepted and promoted.  Evaluated on the device with torch.linalg (tiny matrices,
    off the hot path); also batched: [B,d,d] x [B,d,d] -> [B]."""
    import torch

    da, host_a = _to_device(a)
    db, host_b = _to_device(b)

    def as_dm(t):
        if t.dim() == 1 or (t.dim() == 2 and t.shape[-1] != t.shape[-2]):
            return t[..., :, None] * t.conj()[..., None, :]
        return t

    da, db = as_dm(da), as_dm(db)
    tr_ab = torch.einsum("...ij,...ji->...", da, db).real
    pure = (torch.einsum("...ij,...ji->...", da, da).real > 1 - 1e-12) | (torch.einsum("...ij,...ji->...", db, db).real > 1 - 1e-12)
    if bool(pure.all()):  # a pure argument: F = tr(a b), exact (the general formula loses ~1e-8 to the square roots)
        f = tr_ab
        if host_a and host_b:
            f = f.cpu().numpy()
            return float(f) if f.ndim == 0 else f
        return f
    w, v = torch.linalg.eigh(da)
    root = (v * torch.sqrt(torch.clamp(w, min=0.0)).to(v.dtype)[..., None, :]) @ v.conj().transpose(-1, -2)
    ev = torch.linalg.eigvalsh(root @ db @ root)
    f = torch.sqrt(torch.clamp(ev, min=0.0)).sum(dim=-1) ** 2
    f = torch.where(pure, tr_ab, f)
    if host_a and host_b:
        f = f.cpu().numpy()
        return float(f) if f.ndim == 0 else f
    return f
