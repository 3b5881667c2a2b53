"""Schedules and CSV schema of the training loop against the reference formulas
(experiments/src/utils.jl:1-68, experiments/src/logging.jl:102-128); the GPU part runs a few training
iterations and checks that the loss goes down on a fixed synthetic batch."""
import math

import numpy as np
import pytest

import __graft_entry__ as entry


@pytest.fixture(scope="module")
def tr():
    return entry.load_package().trainer


def test_schedules_follow_the_reference_formulas(tr):
    e = tr.ExponentialDecay(100.0, 10.0, 1000)
    assert abs(e(0) - 100) < 1e-9 and abs(e(1000) - 10) < 1e-9 and abs(e(500) - math.sqrt(1000)) < 1e-9
    assert abs(tr.InverseDecay(1e-3, 1e-4)(10000) - 5e-4) < 1e-12
    s = tr.Step(1.0, 0.1, [10, 20])
    assert s(1) == 1.0 and s(11) == 1.0 and abs(s(12) - 0.1) < 1e-12 and abs(s(22) - 0.01) < 1e-12
    assert tr.Step(1.0, 0.5, 5)(7) == 0.5
    assert tr.Constant(0.3)(99) == 0.3
    c = tr.CosineAnneal(1.0, 0.0, 10, restart=True, dampen=2.0)
    assert abs(c(1) - 1.0) < 1e-12 and abs(c(6) - 0.5) < 1e-12 and abs(c(11) - 0.5) < 1e-12      # restart, dampened
    c2 = tr.CosineAnneal(1.0, 0.0, 10)
    assert abs(c2(11) - 0.0) < 1e-12
    assert tr.w_kl_schedule(50) == 0.0 and abs(tr.w_kl_schedule(200) - (1 - 0.99 ** 100)) < 1e-12
    with pytest.raises(ValueError):
        tr.lr_scheduler("linear", 1e-3)
    assert isinstance(tr.lr_scheduler("inverse", 1e-3), tr.InverseDecay)


def test_csv_headers_match_the_reference_schema(tr, tmp_path):
    assert tr.train_csv_header() == ["Step", "Batch Time", "Data Time", "Forward Pass Time", "Backward Pass Time",
                                     "Optimizer Time", "Cross Entropy Loss", "Regularize Value", "Net Loss", "NFE",
                                     "Accuracy (Top 1)", "Accuracy (Top 5)"]
    assert tr.train_csv_header(latent_ode=True)[6:8] == ["Neg Log Likelihood", "KL Divergence"]
    assert tr.train_csv_header(sde=True)[9:11] == ["NFE Drift", "NFE Diffusion"]
    assert tr.eval_csv_header()[:5] == ["Step", "Batch Time", "Data Time", "Forward Pass Time", "Cross Entropy Loss"]
    log = tr.CSVLogger(str(tmp_path / "results_train.csv"), tr.train_csv_header())
    log(*range(12))
    with pytest.raises(ValueError):
        log(1, 2)
    rows = open(tmp_path / "results_train.csv").read().strip().splitlines()
    assert rows[0].startswith("Step,Batch Time") and rows[1] == ",".join(str(i) for i in range(12))


@pytest.mark.gpu
def test_training_iterations_reduce_the_loss(tr, tmp_path):
    import torch
    t = tr.MnistODETrainer(hidden=100, lr=2e-3, scheduler="constant", w_reg_start=1.0, w_reg_end=0.1, total_steps=10,
                           abstol=1e-3, reltol=1e-3, seed=0)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((256, 784), device="cuda", generator=g)
    y = torch.randint(0, 10, (256,), device="cuda", generator=g, dtype=torch.int32)
    log = tr.CSVLogger(str(tmp_path / "results_train.csv"), tr.train_csv_header())
    rows = []
    for _ in range(8):
        rows.append(t.train_step(x, y))
        log(*rows[-1])
    ce = [r[6] for r in rows]
    assert all(np.isfinite(ce)) and ce[-1] < ce[0] - 0.05, ce            # memorising a fixed batch
    assert rows[0][0] == 1 and rows[-1][0] == 8 and rows[-1][9] > 0 and 0.0 <= rows[-1][10] <= rows[-1][11] <= 100.0
    assert max(r[11] for r in rows) > 1.0        # percentages (experiments/src/utils.jl:71-86), not fractions
    assert len(open(tmp_path / "results_train.csv").read().strip().splitlines()) == 9


@pytest.mark.gpu
def test_cifar10_training_iterations_reduce_the_loss(tr):
    """The cifar10 Chain (construct.jl:212-227) at a reduced size, trained on one fixed synthetic batch through the
    C-ABI calls only: the cross-entropy goes down, the BatchNorm running statistics move, NFE is reported."""
    import torch
    t = tr.Cifar10ODETrainer(hidden=16, image=16, lr=3e-3, scheduler="constant", w_reg_start=1.0, w_reg_end=0.1,
                             total_steps=10, abstol=1e-3, reltol=1e-3, seed=0)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((32, 3, 16, 16), device="cuda", generator=g)
    y = torch.randint(0, 10, (32,), device="cuda", generator=g, dtype=torch.int32)
    r0 = t.st["model"]["running"].copy()
    rows = [t.train_step(x, y) for _ in range(8)]
    ce = [r[6] for r in rows]
    assert all(np.isfinite(ce)) and ce[-1] < ce[0] - 0.05, ce
    assert rows[-1][9] > 0 and 0.0 <= rows[-1][10] <= rows[-1][11] <= 100.0
    r1 = t.st["model"]["running"]
    r1 = r1.cpu().numpy() if hasattr(r1, "cpu") else np.asarray(r1)
    assert np.abs(r1 - r0).max() > 1e-3
    assert float((t.bn_state["running"][:8]).abs().max()) > 0


def _optimisers_jl(kind, p, g, s1, s2, lr, a, b, eps, wd, step):
    """numpy restatement of the Optimisers.jl rules experiments/src/construct.jl:104-125 builds (Descent, Momentum,
    Nesterov, Adam, AdaMax) chained with WeightDecay(wd)."""
    p, g = p.astype(np.float32), g.astype(np.float32)
    f = np.float32
    if kind == "descent":
        dx = f(lr) * g
    elif kind == "momentum":
        s1[:] = f(a) * s1 + f(lr) * g
        dx = s1.copy()
    elif kind == "nesterov":
        dx = -(f(a) * f(a)) * s1 + (f(1) + f(a)) * f(lr) * g
        s1[:] = f(a) * s1 - f(lr) * g
    elif kind == "adamax":
        s1[:] = f(a) * s1 + (f(1) - f(a)) * g
        s2[:] = np.maximum(f(b) * s2, np.abs(g))
        dx = (f(lr) / (f(1) - f(a) ** f(step))) * s1 / (s2 + f(eps))
    else:
        s1[:] = f(a) * s1 + (f(1) - f(a)) * g
        s2[:] = f(b) * s2 + (f(1) - f(b)) * g * g
        dx = f(lr) * (s1 / (f(1) - f(a) ** f(step))) / (np.sqrt(s2 / (f(1) - f(b) ** f(step))) + f(eps))
    return p - (dx + f(wd) * p)


def test_optimiser_construction_follows_the_reference(tr):
    assert tr.Optimiser("adamw").kind == "adam" and tr.Optimiser("adamax").kind == "adamax"
    assert tr.Optimiser("sgd").kind == "descent" and tr.Optimiser("sgd", momentum=0.9).kind == "momentum"
    assert tr.Optimiser("sgd", momentum=0.9, nesterov=True).kind == "nesterov"
    with pytest.raises(ValueError):
        tr.Optimiser("rmsprop")


@pytest.mark.gpu
@pytest.mark.parametrize("optimizer,kw", [("adam", {}), ("adamw", dict(weight_decay=1e-2)), ("adamax", {}),
                                          ("sgd", {}), ("sgd", dict(momentum=0.9)),
                                          ("sgd", dict(momentum=0.9, nesterov=True, weight_decay=1e-3))])
def test_optimiser_rules_match_optimisers_jl(tr, optimizer, kw):
    """lrnde_opt_step against the Optimisers.jl update rules over five steps (experiments/physionet/physionet.yml:27
    trains with AdaMax, the others are reachable through cfg.optimizer)."""
    import torch
    pkg = entry.load_package()
    ctx = pkg.default_context(0)
    opt = tr.Optimiser(optimizer, 3e-3, **kw)
    rng = np.random.default_rng(0)
    n = 5000
    p_ref = rng.standard_normal(n).astype(np.float32)
    p = torch.from_numpy(p_ref.copy()).cuda()
    s1, s2 = np.zeros(n, np.float32), np.zeros(n, np.float32)
    a, b = (opt.beta if opt.kind in ("adam", "adamax") else (opt.momentum, 0.0))
    for step in range(1, 6):
        g = rng.standard_normal(n).astype(np.float32)
        p_ref = _optimisers_jl(opt.kind, p_ref, g, s1, s2, 3e-3, a, b, opt.eps, opt.weight_decay, step)
        opt.update(ctx, "p", p, torch.from_numpy(g).cuda(), step)
    err = np.abs(p.cpu().numpy() - p_ref).max()
    assert err < 2e-6, (optimizer, kw, err)
