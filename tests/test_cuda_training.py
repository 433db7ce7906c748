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


def _dataset_circuit(d):
    name, args, kwargs = d["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    for v in d["x_nodes"]:
        gs[v] = mb.Ment("X")
    return gs


def test_fused_dataset_gradient_matches_reference(monkeypatch):
    """mbqc_psr_grad_dataset (both gradient kernels) against the reference's recorded cost, psr and
    fd gradients and Adam trajectory for the tutorial's data-set averaged infidelity."""
    d = G["dataset"]
    gs = _dataset_circuit(d)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    ins, tgts, x = from_cplx(d["inputs"]), from_cplx(d["targets"]), np.asarray(d["x"])
    for kern in ("pairs", "prefix"):
        monkeypatch.setenv("MBQC_GRAD_KERNEL", kern)
        g, c = mb.gradients.psr_gradient_dataset(ps, x, tgts, ins, return_cost=True)
        assert g.shape == x.shape and abs(float(c) - d["cost"]) < 1e-12
        assert np.allclose(g, d["psr"], atol=1e-11, rtol=0)
        gfd = mb.gradients.psr_gradient_dataset(ps, x, tgts, ins, shift=1e-5)
        assert np.allclose(gfd, d["fd"], atol=1e-6, rtol=0)
    monkeypatch.delenv("MBQC_GRAD_KERNEL")
    cost = mb.optimizers.BatchedFidelityCost(ps, tgts, input_states=ins)
    assert abs(cost(x) - d["cost"]) < 1e-12
    assert np.allclose(mb.gradients.get_gradient(cost, x), d["psr"], atol=1e-11, rtol=0)          # fused hook
    assert np.allclose(mb.gradients.get_gradient(cost, x, method="fd"), d["fd"], atol=1e-6, rtol=0)
    got = mb.optimizers.AdamOptimizer(step_size=0.08).optimize(cost, x.copy(), num_iters=4)
    assert np.allclose(got, d["adam_4"], atol=1e-9, rtol=0)
    # device-resident loop, several parameter vectors at once; row 0 is the golden start
    X0 = np.vstack([x, np.random.default_rng(2).uniform(0, 2 * np.pi, (40, len(x)))])
    out, cst = mb.optimizers.adam_optimize_batched(ps, X0, tgts, num_iters=4, step_size=0.08, input_states=ins,
                                                   dataset=True, return_cost=True)
    assert np.allclose(out[0], d["adam_4"], atol=1e-9, rtol=0)
    assert abs(cst[0] - cost(np.asarray(d["adam_4"]))) < 1e-9


def test_fused_dataset_gradient_properties():
    """P vectors x S items in one launch == per-vector calls == mean of per-sample gradients;
    |+> inputs when input_states is None; argument errors."""
    from scipy.stats import unitary_group

    gs = mb.templates.grid_cluster(3, 4)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    T, rng = len(gs.trainable_nodes), np.random.default_rng(9)
    P, S = 37, 300   # P * S > 32768 / 3: exercise both kernel choices below
    X = rng.uniform(0, 2 * np.pi, (P, T))
    ins = np.stack([unitary_group.rvs(8, random_state=s)[:, 0] for s in range(S)])
    tgts = np.stack([unitary_group.rvs(8, random_state=1000 + s)[:, 0] for s in range(S)])
    g, c = mb.gradients.psr_gradient_dataset(ps, X, tgts, ins, return_cost=True)
    assert g.shape == (P, T) and c.shape == (P,)
    for p in (0, 17, 36):
        gp, cp = mb.gradients.psr_gradient_dataset(ps, X[p], tgts, ins, return_cost=True)
        assert np.allclose(gp, g[p], atol=1e-13) and abs(cp - c[p]) < 1e-13
        per = np.stack([mb.gradients.psr_gradient_batched(ps, X[p][None], tgts[s], input_states=ins[s][None])[0]
                        for s in range(0, S, 50)])
        sub = mb.gradients.psr_gradient_dataset(ps, X[p], tgts[::50], ins[::50])
        assert np.allclose(per.mean(axis=0), sub, atol=1e-13)
    big = mb.gradients.psr_gradient_dataset(ps, np.repeat(X, 4, axis=0), tgts, ins)  # 44,400 samples: prefix kernel
    assert np.allclose(big[::4], g, atol=1e-12)
    plus = mb.gradients.psr_gradient_dataset(ps, X[:3], tgts[:5])
    want = np.mean([mb.gradients.psr_gradient_batched(ps, X[:3], tgts[s]) for s in range(5)], axis=0)
    assert np.allclose(plus, want, atol=1e-13)
    with pytest.raises(ValueError):
        mb.gradients.psr_gradient_dataset(ps, X, tgts, ins[:-1])
    with pytest.raises(ValueError):
        mb.gradients.psr_gradient_dataset(ps, X, tgts[:, :4], None)


def test_fused_training_loop_matches_reference_and_per_iteration_loop():
    """mbqc_train_dataset (all iterations from one C call) against the reference trajectories and
    against the per-iteration loop (fused=False), Adam and SGD / momentum / Nesterov."""
    s, d = G["small"], G["dataset"]
    name, args, kwargs = s["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    ps = mb.PatternSimulator(gs, backend="cuda-sv")
    tgt, x0 = from_cplx(s["target"]), np.asarray(s["x"])
    X0 = np.vstack([x0, np.random.default_rng(1).uniform(0, 2 * np.pi, (30, len(x0)))])
    a = mb.optimizers.adam_optimize_batched(ps, X0, tgt, num_iters=5, step_size=0.1)
    assert np.allclose(a[0], s["adam_5"], atol=1e-9, rtol=0)
    b = mb.optimizers.adam_optimize_batched(ps, X0, tgt, num_iters=5, step_size=0.1, fused=False)
    assert np.allclose(a, b, atol=1e-12, rtol=0)
    for kw, key in (({"momentum": 0.9}, "sgd_mom_5"), ({"momentum": 0.9, "nesterov": True}, "sgd_nesterov_5")):
        f = mb.optimizers.sgd_optimize_batched(ps, X0, tgt, num_iters=5, step_size=0.2, **kw)
        assert np.allclose(f[0], s[key], atol=1e-9, rtol=0)
        u = mb.optimizers.sgd_optimize_batched(ps, X0, tgt, num_iters=5, step_size=0.2, fused=False, **kw)
        assert np.allclose(f, u, atol=1e-12, rtol=0)
    # data-set cost with the training curve
    gd = _dataset_circuit(d)
    pd = mb.PatternSimulator(gd, backend="cuda-sv")
    ins, tgts, x = from_cplx(d["inputs"]), from_cplx(d["targets"]), np.asarray(d["x"])
    out, cost, hist = mb.optimizers.adam_optimize_batched(pd, x[None], tgts, num_iters=4, step_size=0.08, input_states=ins,
                                                          dataset=True, return_cost=True, return_history=True)
    assert np.allclose(out[0], d["adam_4"], atol=1e-9, rtol=0)
    assert hist.shape == (4, 1) and abs(hist[0, 0] - d["cost"]) < 1e-12
    c = mb.optimizers.BatchedFidelityCost(pd, tgts, input_states=ins)
    assert abs(cost[0] - c(out[0])) < 1e-12
    long = mb.optimizers.adam_optimize_batched(pd, x[None], tgts, num_iters=150, step_size=0.08, input_states=ins,
                                               dataset=True, return_history=True)[1]
    assert long[-1, 0] < long[0, 0] - 0.1 and np.all(np.isfinite(long))   # the cost goes down
    # CUDA tensors in -> CUDA tensors out, nothing synchronises
    import torch

    xt = mb.optimizers.adam_optimize_batched(pd, torch.as_tensor(x[None]).cuda(), torch.as_tensor(tgts).cuda(),
                                             num_iters=4, step_size=0.08, input_states=torch.as_tensor(ins).cuda(), dataset=True)
    assert xt.is_cuda and np.allclose(xt.cpu().numpy()[0], d["adam_4"], atol=1e-9, rtol=0)


def test_tutorial_training_script_runs_unchanged():
    """docs/tutorials/intro-to-mbqml.rst:16-86 with `mp` -> `mb` and backend='cuda-sv': the
    reference-style closures (prediction / loss / cost, AdamOptimizer with a callback) work as
    written, and the fused device loop reaches the same parameters."""
    gs = mb.templates.muta(2, 1, one_column=True)
    gs[3] = mb.Ment("X")
    gs[8] = mb.Ment("X")
    ps = mb.PatternSimulator(gs, backend="cuda-sv")

    def loss(output, target):
        avg = 0
        for sty, out in zip(target, output):
            avg += 1 - mb.calculator.fidelity(mb.calculator.pure2density(sty), out)
        return avg / len(target)

    def prediction(thetas, statesx):
        outs = []
        for st in statesx:
            ps.reset(input_state=st)
            outs.append(ps(thetas))
        return outs

    def cost(thetas, statesx, statesy):
        return loss(prediction(thetas, statesx), statesy)

    gate = np.kron(mb.utils.random_special_unitary(1, random_state=3), np.eye(2))
    assert abs(np.linalg.det(gate) - 1) < 1e-12
    (x_train, y_train), (x_test, y_test) = mb.utils.generate_random_dataset(gate, 10, test_size=0.3, random_state=4)
    assert x_train.shape == (7, 4) and x_test.shape == (3, 4) and np.allclose(y_test, x_test @ gate.T)
    curve = []
    theta0 = np.random.default_rng(5).random(len(gs.trainable_nodes))
    opt = mb.optimizers.AdamOptimizer(step_size=0.08)
    theta = opt.optimize(lambda p: cost(p, x_train, y_train), theta0.copy(), num_iters=3,
                         callback=lambda params, it: curve.append(cost(params, x_train, y_train)))
    assert len(curve) == 3 and np.all(np.isfinite(curve))
    fused, hist = mb.optimizers.adam_optimize_batched(ps, theta0[None], y_train, num_iters=3, step_size=0.08,
                                                      input_states=x_train, dataset=True, return_history=True)
    assert np.allclose(fused[0], theta, atol=1e-8)
    assert abs(hist[0, 0] - cost(theta0, x_train, y_train)) < 1e-10
    # fidelity helper: pure/pure, pure/mixed, batched
    a, b = x_train[0], x_train[1]
    assert abs(mb.calculator.fidelity(a, b) - abs(np.vdot(a, b)) ** 2) < 1e-10
    rho = 0.5 * np.outer(a, a.conj()) + 0.5 * np.outer(b, b.conj())
    assert abs(mb.calculator.fidelity(np.outer(a, a.conj()), rho) - np.real(a.conj() @ rho @ a)) < 1e-10
    assert abs(mb.calculator.fidelity(rho, rho) - 1) < 1e-7      # general (Uhlmann) branch
    sig = 0.3 * np.outer(a, a.conj()) + 0.7 * np.eye(4) / 4
    from scipy.linalg import sqrtm
    want = np.real(np.trace(sqrtm(sqrtm(rho) @ sig @ sqrtm(rho)))) ** 2
    assert abs(mb.calculator.fidelity(rho, sig) - want) < 1e-7
