"""Gradient-based optimisers over measurement angles.

Same classes, constructor arguments, update rules and method names as mentpy/optimizers
(adam.py:54-108, sgd.py:49-96, rcd.py:41-99, base_optimizer.py) so existing training loops run
unchanged; every gradient goes through mentpy_b200.gradients.get_gradient, which evaluates all
shifted points in one batched call when the cost exposes `.batch(X)` (see BatchedFidelityCost).
O(T) host arithmetic -- nothing here is accelerated, the cost evaluations are.
"""
import abc
import random

import numpy as np

from ..gradients import get_gradient


class BaseOptimizer(abc.ABC):
    def __init__(self, *args, **kwargs):
        pass

    @abc.abstractmethod
    def step(self, *args, **kwargs):
        pass

    @abc.abstractmethod
    def reset(self, *args, **kwargs):
        pass

    def optimize(self, f, x0, num_iters=100, callback=None, verbose=False, **kwargs):
        x = x0
        for i in range(num_iters):
            x = self.step(f, x, i, **kwargs)
            if callback is not None:
                callback(x, i)
            if verbose:
                print(f"Iteration {i+1}/{num_iters}")
        return x

    def update_step_size(self, x, i, factor=0.99):
        self.step_size = self.step_size * factor


class AdamOptimizer(BaseOptimizer):
    """Adam (Kingma & Ba) with bias-corrected moments; state kept in `m`, `v` across steps."""

    def __init__(self, step_size=0.1, b1=0.9, b2=0.999, eps=10**-8) -> None:
        self.step_size, self.b1, self.b2, self.eps = step_size, b1, b2, eps
        self.m = None
        self.v = None

    def _update(self, g, x, i):
        if self.m is None:
            self.m = np.zeros(len(x))
        if self.v is None:
            self.v = np.zeros(len(x))
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * g**2
        m_hat = self.m / (1 - self.b1 ** (i + 1))
        v_hat = self.v / (1 - self.b2 ** (i + 1))
        return x - self.step_size * m_hat / (np.sqrt(v_hat) + self.eps)

    def step(self, f, x, i, **kwargs):
        return self._update(get_gradient(f, x, **kwargs), x, i)

    def optimize_and_gradient_norm(self, f, x0, num_iters=100, callback=None, verbose=False, **kwargs):
        saved = (self.m, self.v)
        self.m, self.v = np.zeros(len(x0)), np.zeros(len(x0))
        x, norm = x0, np.zeros(num_iters)
        for i in range(num_iters):
            g = get_gradient(f, x, **kwargs)
            x = self._update(g, x, i)
            norm[i] = np.linalg.norm(g)
            if callback is not None:
                callback(x, i)
            if verbose:
                print(f"Iteration {i+1} of {num_iters}: {x} with value {f(x)}")
        self.m, self.v = saved
        return x, norm

    def reset(self):
        self.m = None
        self.v = None


class SGDOptimizer(BaseOptimizer):
    """Gradient descent with optional (Nesterov) momentum."""

    def __init__(self, step_size=0.1, momentum=0.0, nesterov=False) -> None:
        self.step_size, self.momentum, self.nesterov = step_size, momentum, nesterov
        self.v = None

    def _update(self, g, x):
        if self.v is None:
            self.v = np.zeros(len(x))
        self.v = self.momentum * self.v - self.step_size * g
        if self.nesterov:
            return x + self.momentum * self.v - self.step_size * g
        return x + self.v

    def step(self, f, x, i, **kwargs):
        return self._update(get_gradient(f, x, **kwargs), x)

    def optimize_and_gradient_norm(self, f, x0, num_iters=100, callback=None, verbose=False, **kwargs):
        saved, self.v = self.v, np.zeros(len(x0))
        x, norm = x0, []
        for i in range(num_iters):
            g = get_gradient(f, x, **kwargs)
            norm.append(np.linalg.norm(g))
            x = self._update(g, x)
            if callback is not None:
                callback(x, i)
            if verbose:
                print(f"Iteration {i+1} of {num_iters}: {x} with value {f(x)}")
        self.v = saved
        return x, norm

    def reset(self, *args, **kwargs):
        self.v = None


class RCDOptimizer(BaseOptimizer):
    """Random coordinate descent: one central-difference partial derivative (h = 1e-5) per step,
    coordinate drawn with `random.randint`; updates `x` in place like the reference."""

    def __init__(self, step_size=0.1, adaptive=False) -> None:
        self.step_size, self.adaptive = step_size, adaptive

    @staticmethod
    def _partial(f, x, k):
        delta = np.zeros_like(x)
        delta[k] = 1e-5
        if hasattr(f, "batch"):
            fp, fm = f.batch(np.stack([x + delta, x - delta]))
        else:
            fp, fm = f(x + delta), f(x - delta)
        return (fp - fm) / (2 * delta[k])

    def step(self, f, x, i, **kwargs):
        k = random.randint(0, len(x) - 1)
        g = self._partial(f, x, k)
        lr = self.step_size / np.sqrt(i + 1) if self.adaptive else self.step_size
        x[k] -= lr * g
        return x

    def optimize_and_gradient_norm(self, f, x0, num_iters=100, callback=None, verbose=False, **kwargs):
        x, seen, norm = x0, np.zeros(len(x0)), []
        for i in range(num_iters):
            k = random.randint(0, len(x) - 1)
            seen[k] += 1
            g = self._partial(f, x, k)
            lr = self.step_size / np.sqrt(seen[k]) if self.adaptive else self.step_size
            x[k] -= lr * g
            norm.append(np.linalg.norm(g))
            if callback is not None:
                callback(x, i)
            if verbose:
                print(f"Iteration {i+1} of {num_iters}: {x} with value {f(x)}")
        return x, norm

    def reset(self, *args, **kwargs):
        pass


def compute_gradient_variance(f, x, estimate_gradient, num_samples=10, **kwargs):
    """Variance of repeated gradient estimates (mentpy/optimizers/bp_tools.py:10-16)."""
    return np.var(np.array([estimate_gradient(f, x, **kwargs) for _ in range(num_samples)]), axis=0)


class BatchedFidelityCost:
    """cost(x) = mean over the data set of  1 - <t_s| rho_out(x; in_s) |t_s>  -- the training cost
    of docs/tutorials/intro-to-mbqml.rst:35-54 -- as a callable that ALSO exposes `batch(X)`, so
    get_gradient / the optimisers evaluate all 2T shifted angle vectors x all S data states in one
    kernel launch instead of 2*T*S sequential `ps.reset(); ps(x)` calls."""

    def __init__(self, simulator, targets, input_states=None):
        self.sim = getattr(simulator, "simulator", simulator)
        self.targets = np.atleast_2d(np.asarray(targets, dtype=np.complex128))
        self.inputs = None if input_states is None else np.atleast_2d(np.asarray(input_states, dtype=np.complex128))
        if self.inputs is not None and len(self.inputs) != len(self.targets):
            raise ValueError("need one target state per input state")

    def batch(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        n, S = len(X), len(self.targets)
        if self.inputs is None:
            psi = self.sim.run_batch(X, output_form="sv")
            fid = np.abs(psi @ self.targets.conj().T) ** 2          # [n, S]
        else:
            big = np.repeat(X, S, axis=0)
            ins = np.tile(self.inputs, (n, 1))
            psi = self.sim.run_batch(big, input_states=ins, output_form="sv").reshape(n, S, -1)
            fid = np.abs(np.einsum("nsk,sk->ns", psi, self.targets.conj())) ** 2
        return 1.0 - fid.mean(axis=1)

    def __call__(self, x):
        return float(self.batch(np.asarray(x)[None, :])[0])

    def _fused_ok(self):
        from .. import _lib

        plan = getattr(self.sim, "plan", None)
        return hasattr(self.sim, "_full_plan") and plan is not None and not getattr(plan, "mixed", False) \
            and plan.window <= _lib.MAX_WINDOW_REG and getattr(self.sim, "dtype", "complex128") == "complex128"

    def shift_gradient(self, x, shift=1.5):
        """(cost(x + s e_i) - cost(x - s e_i)) / (2 s) for every i -- the reference's psr / central
        fd formula -- from the fused data-set kernel; falls back to `batch` on other backends."""
        if not self._fused_ok():
            x = np.asarray(x, dtype=float)
            eye = np.eye(len(x))
            v = self.batch(np.concatenate([x + shift * eye, x - shift * eye]))
            return (v[: len(x)] - v[len(x):]) / (2 * shift)
        from ..gradients import psr_gradient_dataset

        return psr_gradient_dataset(self.sim, x, self.targets, self.inputs, shift=shift)
