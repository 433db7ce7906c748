"""Device-resident training loop: B independent angle vectors optimised in parallel.  Default
(`fused=True`): ALL iterations are queued by one C call (mbqc_train_dataset) -- per iteration the
fused gradient kernel and a reduce kernel that also applies the Adam / SGD update; fallback
(one input state per parameter vector, or fused=False): one fused gradient launch
(mbqc_psr_grad_batch) + elementwise torch updates per iteration.  Nothing crosses PCIe between
iterations either way.

The update rules are the reference's (mentpy/optimizers/adam.py:54-66, sgd.py:49-59) applied row
by row, and the gradient is its shift-1.5 central difference (gradients/_parameter_shift.py:20-24),
so row b of the result equals `AdamOptimizer().optimize(cost_b, x0_b, num_iters)` on the host
(tests/test_cuda_training.py pins this against the reference trajectory in tests/golden/).
"""
from typing import Optional

import numpy as np


def _grad(sim, X, target, shift, input_states, dataset=False, return_cost=False):
    from ..gradients import psr_gradient_batched, psr_gradient_dataset

    if dataset:  # target = [S, 2^k] targets, input_states = [S, 2^|I|] or None; mean over the data set
        return psr_gradient_dataset(sim, X, target, input_states, shift=shift, return_cost=return_cost)
    return psr_gradient_batched(sim, X, target, shift=shift, input_states=input_states, return_cost=return_cost)


def _fused_train(sim, x0, target, input_states, dataset, kind, num_iters, shift, return_cost, return_history=False, **opt):
    """All iterations in ONE C call (mbqc_train_dataset): per iteration the fused data-set gradient
    kernel + a reduce kernel that also applies the optimiser update; no Python, no host round trip
    between iterations.  Returns None when the configuration is not covered (caller falls back
    to the per-iteration loop)."""
    import ctypes as C

    import torch

    from .. import _lib

    plan = getattr(sim, "plan", None)
    if plan is None or getattr(plan, "mixed", False) or plan.window > _lib.MAX_WINDOW_REG \
            or getattr(sim, "dtype", "complex128") != "complex128" or num_iters <= 0:
        return None
    dev = sim._dev()
    dplan = sim._full_plan()
    on_host = not isinstance(x0, torch.Tensor)
    X = torch.as_tensor(np.atleast_2d(x0) if on_host else x0, dtype=torch.float64).to(dev).clone()
    if X.dim() == 1:
        X = X[None, :]
    X = X.contiguous()
    P, T = X.shape
    if T != plan.n_angles:
        raise ValueError(f"Number of angles ({T}) does not match number of trainable nodes ({plan.n_angles}).")

    def stage(a, width):
        t = a.to(device=dev, dtype=torch.complex128) if isinstance(a, torch.Tensor) \
            else torch.as_tensor(np.ascontiguousarray(a, dtype=np.complex128)).to(dev)
        return t.reshape(-1, t.shape[-1]).contiguous() if t.shape[-1] == width else None

    tg = stage(target, 2 ** dplan.n_out)
    if tg is None:
        raise ValueError(f"target must have {2 ** dplan.n_out} amplitudes per state")
    if not dataset and tg.shape[0] != 1:
        return None
    inp = None
    if input_states is not None:
        inp = stage(input_states, 2 ** dplan.n_in)
        if inp is None:
            raise ValueError(f"input_states must have {2 ** dplan.n_in} amplitudes per state")
        if not dataset and inp.shape[0] != 1:
            return None  # one input per parameter vector: not a data-set layout
    elif sim.input_state is not None:
        inp = stage(sim.input_state, 2 ** dplan.n_in)
    S = tg.shape[0]
    if inp is not None and inp.shape[0] != S:
        if inp.shape[0] == 1:
            inp = inp.expand(S, -1).contiguous()
        else:
            raise ValueError("need one target state per input state")
    lib = _lib.load()
    with torch.cuda.device(dev):
        state = torch.zeros((2, P, T), dtype=torch.float64, device=dev)
        hist = torch.empty((num_iters, P), dtype=torch.float64, device=dev) if return_history else None
        status = torch.zeros(P, dtype=torch.int32, device=dev)
        ws = torch.empty(max(int(lib.mbqc_psr_grad_dataset_workspace_bytes(dplan.handle, P, S)), 16), dtype=torch.uint8, device=dev)
        o = _lib.Optimizer(kind=kind, nesterov=int(bool(opt.get("nesterov", False))), step_size=opt["step_size"],
                           b1=opt.get("b1", 0.9), b2=opt.get("b2", 0.999), eps=opt.get("eps", 1e-8),
                           momentum=opt.get("momentum", 0.0))
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mbqc_train_dataset(dplan.handle, X.data_ptr(), None if inp is None else inp.data_ptr(), tg.data_ptr(),
                                          P, S, C.c_double(shift), C.byref(o), 0, num_iters, state.data_ptr(),
                                          None if hist is None else hist.data_ptr(), status.data_ptr(), ws.data_ptr(), stream))
        cost = None
        if return_cost:  # cost at the optimised point: one more gradient-free evaluation of the loop's cost
            from ..gradients import psr_gradient_dataset

            _, cost = psr_gradient_dataset(sim, X, tg, inp, shift=shift, return_cost=True)
        if on_host:
            sim._check_status(status)
            return (X.cpu().numpy(), None if cost is None else cost.cpu().numpy(),
                    None if hist is None else hist.cpu().numpy())
        sim.last_status = status
        return X, cost, hist


def adam_optimize_batched(simulator, x0, target, num_iters: int = 100, step_size: float = 0.1,
                          b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, shift: float = 1.5,
                          input_states=None, return_cost: bool = False, dataset: bool = False,
                          fused: bool = True, return_history: bool = False):
    """Adam on every row of x0 [B,T] for the cost 1 - |<target|psi_out(x)>|^2.  Returns the
    optimised angles as a numpy array (numpy in) or CUDA tensor (tensor in); with return_cost also
    the cost at the optimised point; with return_history also the cost before every update
    [num_iters, B] (the training curve the tutorial records through its callback).

    dataset=True: `target` is [S,2^k] and `input_states` [S,2^|I|] (or None); every row of x0 is
    trained on the data-set averaged cost (docs/tutorials/intro-to-mbqml.rst:35-86), one fused
    mbqc_psr_grad_dataset launch per iteration."""
    import torch

    sim = getattr(simulator, "simulator", simulator)
    if fused:
        res = _fused_train(sim, x0, target, input_states, dataset, 1, num_iters, shift, return_cost, return_history,
                           step_size=step_size, b1=b1, b2=b2, eps=eps)
        if res is not None:
            out = (res[0],) + ((res[1],) if return_cost else ()) + ((res[2],) if return_history else ())
            return out if len(out) > 1 else out[0]
    if return_history:
        raise NotImplementedError("return_history needs the fused loop (shared or data-set inputs, window <= 5)")
    dev = sim._dev()
    on_host = not isinstance(x0, torch.Tensor)
    X = torch.as_tensor(np.atleast_2d(x0) if on_host else x0, dtype=torch.float64).to(dev).clone()
    if X.dim() == 1:
        X = X[None, :]
    m = torch.zeros_like(X)
    v = torch.zeros_like(X)
    for i in range(num_iters):
        g = _grad(sim, X, target, shift, input_states, dataset)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        m_hat = m / (1 - b1 ** (i + 1))
        v_hat = v / (1 - b2 ** (i + 1))
        X = X - step_size * m_hat / (torch.sqrt(v_hat) + eps)
    out = X.cpu().numpy() if on_host else X
    if return_cost:
        _, c = _grad(sim, X, target, shift, input_states, dataset, return_cost=True)
        return out, (c.cpu().numpy() if on_host else c)
    return out


def sgd_optimize_batched(simulator, x0, target, num_iters: int = 100, step_size: float = 0.1,
                         momentum: float = 0.0, nesterov: bool = False, shift: float = 1.5,
                         input_states=None, dataset: bool = False, fused: bool = True):
    """SGD (+momentum / Nesterov) on every row of x0, same update as the reference's SGDOptimizer
    (dataset=True: data-set averaged cost, see adam_optimize_batched)."""
    import torch

    sim = getattr(simulator, "simulator", simulator)
    if fused:
        res = _fused_train(sim, x0, target, input_states, dataset, 2, num_iters, shift, False,
                           step_size=step_size, momentum=momentum, nesterov=nesterov)
        if res is not None:
            return res[0]
    dev = sim._dev()
    on_host = not isinstance(x0, torch.Tensor)
    X = torch.as_tensor(np.atleast_2d(x0) if on_host else x0, dtype=torch.float64).to(dev).clone()
    vel = torch.zeros_like(X)
    for _ in range(num_iters):
        g = _grad(sim, X, target, shift, input_states, dataset)
        vel = momentum * vel - step_size * g
        X = X + momentum * vel - step_size * g if nesterov else X + vel
    return X.cpu().numpy() if on_host else X
