"""Parameter-shift and finite-difference derivatives.

Same call signatures and the same formulas as mentpy/gradients/grad.py:13-56,
_parameter_shift.py:9-48 and _finite_difference.py:9-58.  NB: the reference's "parameter-shift"
gradient is literally a central difference with step `shift` = 1.5 divided by 2*shift (no sine
factor); that formula is kept verbatim.

Acceleration hooks (additive):
  * a cost object exposing `batch(X) -> costs` (X is [n, T]) gets all its shifted evaluations in
    ONE call instead of 2T (or 4T^2) sequential ones -- see `BatchedCost` users in optimizers;
  * a cost object exposing `shift_gradient(x, shift)` (BatchedFidelityCost on a CUDA SV backend)
    gets the whole central difference from the fused data-set kernel (mbqc_psr_grad_dataset);
  * `psr_gradient_batched` runs B gradients of the fidelity cost 1 - |<t|psi(x)>|^2 entirely in
    the fused CUDA kernel (mbqc_psr_grad_batch), never materialising shifted angle vectors.
"""
import ctypes as C

import numpy as np


def _evaluate_many(cost, points):
    """cost at every row of `points` -- one batched call when the cost supports it."""
    if hasattr(cost, "batch"):
        return np.asarray(cost.batch(np.asarray(points)), dtype=float)
    return np.array([cost(p) for p in points], dtype=float)


def psr_gradient(cost, x, shift=1.5):
    if hasattr(cost, "shift_gradient"):  # fused kernel: all 2T x S evaluations in one launch
        return np.asarray(cost.shift_gradient(x, shift), dtype=float)
    x = np.asarray(x, dtype=float)
    n = len(x)
    eye = np.eye(n)
    vals = _evaluate_many(cost, np.concatenate([x + shift * eye, x - shift * eye]))
    return (vals[:n] - vals[n:]) / (2 * shift)


def psr_hessian(cost, x, shift=1.5):
    x = np.asarray(x, dtype=float)
    n = len(x)
    eye = np.eye(n)
    pts = []
    for i in range(n):
        for j in range(n):
            for si, sj in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                pts.append(x + si * shift * eye[i] + sj * shift * eye[j])
    v = _evaluate_many(cost, np.array(pts)).reshape(n, n, 4)
    return (v[..., 0] - v[..., 1] - v[..., 2] + v[..., 3]) / (4 * shift**2)


def fd_gradient(f, x, h=1e-5, type="central"):
    if type not in ["central", "forward", "backward"]:
        raise UserWarning(f"Expected type to be 'central', 'forward', or 'backward' but {type} was given")
    x = np.asarray(x, dtype=float)
    n = len(x)
    eye = np.eye(n)
    if type == "central":
        if hasattr(f, "shift_gradient"):
            return np.asarray(f.shift_gradient(x, h), dtype=float)
        v = _evaluate_many(f, np.concatenate([x + h * eye, x - h * eye]))
        return (v[:n] - v[n:]) / (2 * h)
    if type == "forward":
        v = _evaluate_many(f, np.concatenate([x + h * eye, x[None, :]]))
        return (v[:n] - v[n]) / h
    v = _evaluate_many(f, np.concatenate([x[None, :], x - h * eye]))
    return (v[0] - v[1:]) / h


def fd_hessian(f, x, h=1e-5, type="central"):
    if type not in ["central", "forward", "backward"]:
        raise UserWarning(f"Expected type to be 'central', 'forward', or 'backward' but {type} was given")
    x = np.asarray(x, dtype=float)
    n = len(x)
    eye = np.eye(n)
    if type == "central":
        return psr_hessian(f, x, shift=h)
    sgn = 1.0 if type == "forward" else -1.0
    pts = [x]
    for i in range(n):
        pts.append(x + sgn * h * eye[i])
    for i in range(n):
        for j in range(n):
            pts.append(x + sgn * h * eye[i] + sgn * h * eye[j])
    v = _evaluate_many(f, np.array(pts))
    f0, fi, fij = v[0], v[1 : n + 1], v[n + 1 :].reshape(n, n)
    return (fij - fi[:, None] - fi[None, :] + f0) / h**2


def get_gradient(cost, x, method="parameter-shift", *args, **kwargs):
    if method in ("parameter-shift", "psr", "parametershift"):
        return psr_gradient(cost, x, *args, **kwargs)
    if method in ("finite-differences", "fd", "finitedifferences"):
        return fd_gradient(cost, x, *args, **kwargs)
    raise UserWarning(f"Expected method to be 'parameter-shift' or 'finite-difference' but {method} was given")


def get_hessian(cost, x, method="parameter-shift", *args, **kwargs):
    if method in ("parameter-shift", "psr", "parametershift"):
        return psr_hessian(cost, x, *args, **kwargs)
    if method in ("finite-differences", "fd", "finitedifferences"):
        return fd_hessian(cost, x, *args, **kwargs)
    raise UserWarning(f"Expected method to be 'parameter-shift' or 'finite-difference' but {method} was given")


def psr_gradient_batched(simulator, angles, target, shift=1.5, input_states=None, return_cost=False):
    """B gradients of cost(x) = 1 - |<target|psi_out(x)>|^2 in one fused kernel launch.

    simulator: PatternSimulator / CudaSimulatorSV (window <= 5); angles [B,T] numpy or torch CUDA;
    target [2^k].  Returns grad [B,T] (and cost [B]) as numpy (numpy in) or torch (torch in)."""
    import torch

    from .. import _lib

    sim = getattr(simulator, "simulator", simulator)
    dev = sim._dev()
    lib = _lib.load()
    with torch.cuda.device(dev):
        a, on_host = sim._stage_angles(angles, dev)
        batch, T = a.shape
        inp, mode = sim._stage_inputs(input_states, batch, dev)
        dplan = sim._full_plan()
        tgt = torch.as_tensor(np.ascontiguousarray(target, dtype=np.complex128)).to(dev) \
            if not isinstance(target, torch.Tensor) else target.to(device=dev, dtype=torch.complex128).contiguous()
        if tgt.numel() != 2 ** dplan.n_out:
            raise ValueError(f"target must have {2 ** dplan.n_out} amplitudes")
        grad = torch.empty((batch, T), dtype=torch.float64, device=dev)
        cost = torch.empty(batch, dtype=torch.float64, device=dev) if return_cost else None
        status = torch.empty(batch, dtype=torch.int32, device=dev)
        _lib.check(lib.mbqc_psr_grad_batch(dplan.handle, a.data_ptr(), (a.stride(0) if batch > 1 else max(T, 1)),
                                           None if inp is None else inp.data_ptr(), mode, batch,
                                           tgt.data_ptr(), C.c_double(shift), grad.data_ptr(),
                                           None if cost is None else cost.data_ptr(),
                                           status.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream))
        if on_host:
            g = grad.cpu().numpy()
            sim._check_status(status)
            return (g, cost.cpu().numpy()) if return_cost else g
        sim.last_status = status
        return (grad, cost) if return_cost else grad


def psr_gradient_dataset(simulator, angles, targets, input_states=None, shift=1.5, return_cost=False):
    """Gradient of the data-set averaged cost  mean_s [1 - |<t_s|psi_out(x; in_s)>|^2]  for P angle
    vectors in one fused launch (mbqc_psr_grad_dataset): the S x 2T `ps.reset(); ps(x)` calls that
    one optimiser step of docs/tutorials/intro-to-mbqml.rst:35-86 makes per parameter vector.

    angles [P,T] (or [T]) numpy / torch CUDA; targets [S,2^k]; input_states [S,2^|I|] or None
    (|+> inputs, S = len(targets)).  Returns grad [P,T] (and cost [P]), numpy in -> numpy out."""
    import torch

    from .. import _lib

    sim = getattr(simulator, "simulator", simulator)
    dev = sim._dev()
    lib = _lib.load()
    with torch.cuda.device(dev):
        squeeze = (angles.dim() if isinstance(angles, torch.Tensor) else np.ndim(angles)) == 1
        a, on_host = sim._stage_angles(angles, dev)
        P, T = a.shape
        dplan = sim._full_plan()

        def stage(x, width, what):
            t = x.to(device=dev, dtype=torch.complex128) if isinstance(x, torch.Tensor) \
                else torch.as_tensor(np.ascontiguousarray(np.atleast_2d(x), dtype=np.complex128)).to(dev)
            t = t.reshape(-1, t.shape[-1]).contiguous()
            if t.shape[1] != width:
                raise ValueError(f"{what} must have {width} amplitudes per state (got {t.shape[1]})")
            return t

        tg = stage(targets, 2 ** dplan.n_out, "targets")
        S = tg.shape[0]
        if input_states is None and not sim._input_is_plus():
            # the cost is evaluated with the simulator's own input state (run_batch: INPUT_SHARED):
            # differentiate that cost, not the |+> one
            input_states = np.tile(np.asarray(sim.input_state, dtype=np.complex128), (S, 1))
        inp = None if input_states is None else stage(input_states, 2 ** dplan.n_in, "input_states")
        if inp is not None and inp.shape[0] != S:
            raise ValueError("need one target state per input state")
        grad = torch.empty((P, T), dtype=torch.float64, device=dev)
        cost = torch.empty(P, dtype=torch.float64, device=dev)
        status = torch.empty(P, dtype=torch.int32, device=dev)
        nbytes = lib.mbqc_psr_grad_dataset_workspace_bytes(dplan.handle, P, S)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
        _lib.check(lib.mbqc_psr_grad_dataset(dplan.handle, a.data_ptr(), (a.stride(0) if P > 1 else max(T, 1)),
                                             None if inp is None else inp.data_ptr(), tg.data_ptr(), P, S,
                                             C.c_double(shift), grad.data_ptr(), cost.data_ptr(),
                                             status.data_ptr(), ws.data_ptr(),
                                             torch.cuda.current_stream(dev).cuda_stream))
        if squeeze:
            grad, cost = grad[0], cost[0]
        if on_host:
            sim._check_status(status)
            g, c = grad.cpu().numpy(), cost.cpu().numpy()
            return (g, c) if return_cost else g
        sim.last_status = status
        return (grad, cost) if return_cost else grad
