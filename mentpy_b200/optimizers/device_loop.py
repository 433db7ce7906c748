"""Device-resident training loop: B independent angle vectors optimised in parallel, every
iteration = ONE fused gradient launch (mbqc_psr_grad_batch) + an elementwise update on the GPU.
Nothing crosses PCIe between iterations.

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


def adam_optimize_batched(simulator, x0, target, num_iters: int = 100, step_size: float = 0.1,
                          b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, shift: float = 1.5,
                          input_states=None, return_cost: bool = False, dataset: bool = False):
    """Adam on every row of x0 [B,T] for the cost 1 - |<target|psi_out(x)>|^2.  Returns the
    optimised angles as a numpy array (numpy in) or CUDA tensor (tensor in).

    dataset=True: `target` is [S,2^k] and `input_states` [S,2^|I|] (or None); every row of x0 is
    trained on the data-set averaged cost (docs/tutorials/intro-to-mbqml.rst:35-86), one fused
    mbqc_psr_grad_dataset launch per iteration."""
    import torch

    sim = getattr(simulator, "simulator", simulator)
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
                         input_states=None, dataset: bool = False):
    """SGD (+momentum / Nesterov) on every row of x0, same update as the reference's SGDOptimizer
    (dataset=True: data-set averaged cost, see adam_optimize_batched)."""
    import torch

    sim = getattr(simulator, "simulator", simulator)
    dev = sim._dev()
    on_host = not isinstance(x0, torch.Tensor)
    X = torch.as_tensor(np.atleast_2d(x0) if on_host else x0, dtype=torch.float64).to(dev).clone()
    vel = torch.zeros_like(X)
    for _ in range(num_iters):
        g = _grad(sim, X, target, shift, input_states, dataset)
        vel = momentum * vel - step_size * g
        X = X + momentum * vel - step_size * g if nesterov else X + vel
    return X.cpu().numpy() if on_host else X
