"""GPU: the reference's training loop (get_gradient -> optimiser step) on CUDA-evaluated costs
reproduces trajectories recorded from the unmodified reference (tests/golden/gradients.json)."""
import numpy as np
import pytest

import mentpy_b200 as mb
from conftest import from_cplx, load_golden

pytestmark = pytest.mark.gpu
G = load_golden("gradients.json")


def test_adam_sgd_trajectories_on_cuda_costs():
    s = G["small"]
    name, args, kwargs = s["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    tgt = from_cplx(s["target"])
    cost = mb.optimizers.BatchedFidelityCost(ps, tgt)
    x0 = np.asarray(s["x"])
    assert abs(cost(x0) - s["cost"]) < 1e-12
    assert np.allclose(mb.gradients.get_gradient(cost, x0), s["psr"], atol=1e-11, rtol=0)
    assert np.allclose(mb.optimizers.AdamOptimizer(step_size=0.1).optimize(cost, x0.copy(), num_iters=5),
                       s["adam_5"], atol=1e-9, rtol=0)
    assert np.allclose(mb.optimizers.SGDOptimizer(step_size=0.2, momentum=0.9).optimize(cost, x0.copy(), num_iters=5),
                       s["sgd_mom_5"], atol=1e-9, rtol=0)

    # the reference's own closure style (docs/tutorials/intro-to-mbqml.rst:43-54) also works
    def closure(x):
        ps.reset()
        rho = ps.run(x)
        return float(1 - np.real(tgt.conj() @ rho @ tgt))

    assert np.allclose(mb.gradients.get_gradient(closure, x0), s["psr"], atol=1e-11, rtol=0)


def test_dataset_cost_with_input_states():
    """S Haar input/target pairs of a fixed unitary (the QML workload): batched cost == loop."""
    from scipy.stats import unitary_group

    gs = mb.templates.grid_cluster(2, 5)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    U = unitary_group.rvs(4, random_state=1)
    ins = np.stack([unitary_group.rvs(4, random_state=10 + s)[:, 0] for s in range(6)])
    tgts = ins @ U.T
    cost = mb.optimizers.BatchedFidelityCost(ps, tgts, input_states=ins)
    x = np.random.default_rng(0).uniform(0, 2 * np.pi, len(gs.trainable_nodes))
    want = 0.0
    for st, tg in zip(ins, tgts):
        ps.reset(input_state=st)
        rho = ps.run(x)
        want += 1 - np.real(tg.conj() @ rho @ tg)
    assert abs(cost(x) - want / len(ins)) < 1e-12
    g = mb.gradients.get_gradient(cost, x)
    assert g.shape == x.shape and np.all(np.isfinite(g))


def test_device_resident_adam_and_sgd_match_reference_trajectories():
    """B parameter vectors optimised in parallel on the device; row 0 starts at the golden x0 and
    must reproduce the reference's Adam / SGD trajectories (tests/golden/gradients.json)."""
    s = G["small"]
    name, args, kwargs = s["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    tgt = from_cplx(s["target"])
    x0 = np.asarray(s["x"])
    X0 = np.vstack([x0, np.random.default_rng(1).uniform(0, 2 * np.pi, (63, len(x0)))])
    X, cost = mb.optimizers.adam_optimize_batched(ps, X0, tgt, num_iters=5, step_size=0.1, return_cost=True)
    assert X.shape == X0.shape and np.allclose(X[0], s["adam_5"], atol=1e-9, rtol=0)
    assert np.all(np.isfinite(cost))
    Y = mb.optimizers.sgd_optimize_batched(ps, X0, tgt, num_iters=5, step_size=0.2, momentum=0.9)
    assert np.allclose(Y[0], s["sgd_mom_5"], atol=1e-9, rtol=0)
    Z = mb.optimizers.sgd_optimize_batched(ps, X0, tgt, num_iters=5, step_size=0.2, momentum=0.9, nesterov=True)
    assert np.allclose(Z[0], s["sgd_nesterov_5"], atol=1e-9, rtol=0)
    # every row equals the host optimiser run on that row alone
    cost_fn = mb.optimizers.BatchedFidelityCost(ps, tgt)
    ref = mb.optimizers.AdamOptimizer(step_size=0.1).optimize(cost_fn, X0[7].copy(), num_iters=5)
    assert np.allclose(X[7], ref, atol=1e-9, rtol=0)
    # longer run actually trains: mean cost decreases
    _, c0 = mb.optimizers.adam_optimize_batched(ps, X0, tgt, num_iters=0, return_cost=True)
    _, c1 = mb.optimizers.adam_optimize_batched(ps, X0, tgt, num_iters=40, return_cost=True)
    assert c1.mean() < c0.mean() - 0.1
