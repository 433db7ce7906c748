"""CPU: gradient formulas and optimiser update rules against values produced by the unmodified
reference (tests/golden/gradients.json).  The cost here is evaluated with the numpy oracle so the
test needs no GPU; the CUDA-evaluated twins are in test_cuda_parity.py / test_cuda_training.py."""
import random

import numpy as np
import pytest

import mentpy_b200 as mb
from conftest import from_cplx, load_golden
from oracle import matrix_free
from oracle.pattern_data import PatternData

G = load_golden("gradients.json")


class OracleCost:
    def __init__(self, spec, target):
        name, args, kwargs = spec
        self.pat = PatternData.from_circuit(getattr(mb.templates, name)(*args, **kwargs))
        self.t = target
        self.batched_calls = 0

    def batch(self, X):
        self.batched_calls += 1
        psi = matrix_free.run_sv_batch(self.pat, np.atleast_2d(X))
        return 1 - np.abs(psi @ self.t.conj()) ** 2

    def __call__(self, x):
        return float(self.batch(np.asarray(x)[None, :])[0])


def test_psr_and_fd_gradient_match_reference():
    g = G["c4"]
    cost = OracleCost(g["spec"], from_cplx(g["target"]))
    x = np.asarray(g["x"])
    assert abs(cost(x) - g["cost"]) < 1e-12
    assert np.allclose(mb.gradients.get_gradient(cost, x), g["psr"], atol=1e-11, rtol=0)
    assert cost.batched_calls == 2  # cost(x) + ONE batched call for all 2T shifts
    assert np.allclose(mb.gradients.get_gradient(cost, x, method="fd"), g["fd"], atol=1e-6, rtol=0)
    plain = lambda v: cost(v)  # no .batch attribute: sequential path, same numbers
    assert np.allclose(mb.gradients.psr_gradient(plain, x), g["psr"], atol=1e-11, rtol=0)
    with pytest.raises(UserWarning):
        mb.gradients.get_gradient(cost, x, method="nope")
    with pytest.raises(UserWarning):
        mb.gradients.fd_gradient(cost, x, type="sideways")
    fwd = mb.gradients.fd_gradient(cost, x, h=1e-6, type="forward")
    bwd = mb.gradients.fd_gradient(cost, x, h=1e-6, type="backward")
    assert np.allclose(fwd, g["fd"], atol=1e-4) and np.allclose(bwd, g["fd"], atol=1e-4)


def test_optimizers_match_reference_trajectories():
    s = G["small"]
    cost = OracleCost(s["spec"], from_cplx(s["target"]))
    x0 = np.asarray(s["x"])
    assert abs(cost(x0) - s["cost"]) < 1e-12
    assert np.allclose(mb.gradients.get_gradient(cost, x0), s["psr"], atol=1e-11, rtol=0)
    assert np.allclose(mb.gradients.get_hessian(cost, x0)[0, :2], s["hessian_psr_00_01"], atol=1e-10)
    adam = mb.optimizers.AdamOptimizer(step_size=0.1)
    assert np.allclose(adam.optimize(cost, x0.copy(), num_iters=5), s["adam_5"], atol=1e-9, rtol=0)
    sgd = mb.optimizers.SGDOptimizer(step_size=0.2, momentum=0.9)
    assert np.allclose(sgd.optimize(cost, x0.copy(), num_iters=5), s["sgd_mom_5"], atol=1e-9, rtol=0)
    nes = mb.optimizers.SGDOptimizer(step_size=0.2, momentum=0.9, nesterov=True)
    assert np.allclose(nes.optimize(cost, x0.copy(), num_iters=5), s["sgd_nesterov_5"], atol=1e-9, rtol=0)
    random.seed(7)
    rcd = mb.optimizers.RCDOptimizer(step_size=0.3, adaptive=True)
    assert np.allclose(rcd.optimize(cost, x0.copy(), num_iters=6), s["rcd_seed7_6"], atol=1e-6, rtol=0)
    adam.reset(); sgd.reset(); rcd.reset()
    assert adam.m is None and sgd.v is None
    x, norms = mb.optimizers.AdamOptimizer().optimize_and_gradient_norm(cost, x0.copy(), num_iters=2)
    assert len(norms) == 2 and x.shape == x0.shape
    var = mb.optimizers.compute_gradient_variance(cost, x0, mb.gradients.get_gradient, num_samples=2)
    assert np.allclose(var, 0)


def _dataset_circuit(d):
    name, args, kwargs = d["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    for v in d["x_nodes"]:
        gs[v] = mb.Ment("X")
    return gs


def test_dataset_averaged_cost_matches_reference():
    """The tutorial's data-set averaged infidelity (docs/tutorials/intro-to-mbqml.rst:35-54) and its
    psr / fd gradients + 4 Adam steps, oracle-evaluated, against the reference's recorded values."""
    d = G["dataset"]
    pat = PatternData.from_circuit(_dataset_circuit(d))
    ins, tgts = from_cplx(d["inputs"]), from_cplx(d["targets"])
    S = len(ins)

    class Cost:
        def batch(self, X):
            X = np.atleast_2d(X)
            psi = matrix_free.run_sv_batch(pat, np.repeat(X, S, axis=0), input_states=np.tile(ins, (len(X), 1)))
            fid = np.abs(np.einsum("nsk,sk->ns", psi.reshape(len(X), S, -1), tgts.conj())) ** 2
            return 1 - fid.mean(axis=1)

        def __call__(self, x):
            return float(self.batch(np.asarray(x)[None])[0])

    cost, x = Cost(), np.asarray(d["x"])
    assert abs(cost(x) - d["cost"]) < 1e-12
    assert np.allclose(mb.gradients.get_gradient(cost, x), d["psr"], atol=1e-11, rtol=0)
    assert np.allclose(mb.gradients.get_gradient(cost, x, method="fd"), d["fd"], atol=1e-6, rtol=0)
    got = mb.optimizers.AdamOptimizer(step_size=0.08).optimize(cost, x.copy(), num_iters=4)
    assert np.allclose(got, d["adam_4"], atol=1e-9, rtol=0)
