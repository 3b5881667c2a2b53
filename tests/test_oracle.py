"""Self-validation of the CPU oracle (parity is unpinned by the reference: its tests
hold no golden vectors, test/runtests.jl:21-29,118-131).  Everything here runs on CPU."""
import copy

import numpy as np
import pytest

from oracle import (MLP, Dense, NeuralODE, Tableau, adjoint_backward, fastpow,
                    glorot_uniform_params, ode_initdt, perform_step_sosri_reg,
                    reg_step_backward, solve_tsit5, SOSRI, diffeqsol_to_array,
                    diffeqsol_to_timeseries, perform_step_tsit5_reg)
from oracle.lrnde_oracle import tsit5_stages, interp_step


def test_tableau_order_conditions():
    tab = Tableau(np.float64)
    b = np.array(tab.a[6] + [0.0])
    c = np.array(tab.c)
    A = np.zeros((7, 7))
    for i, row in enumerate(tab.a):
        A[i, :len(row)] = row
    assert np.allclose(A.sum(1), c, atol=1e-14)
    assert abs(b.sum() - 1) < 1e-14
    for p in range(1, 5):                       # quadrature conditions to order 5
        assert abs(b @ c ** p - 1 / (p + 1)) < 1e-13
    assert abs(b @ (A @ c) - 1 / 6) < 1e-13
    assert abs(b @ (A @ (A @ c)) - 1 / 24) < 1e-13
    assert abs(b @ (A @ (A @ (A @ c))) - 1 / 120) < 1e-13
    assert abs(sum(tab.btilde)) < 1e-14         # embedded pair shares order-1 condition
    w = tab.interp_weights(1.0)
    assert np.allclose(w, b, atol=1e-13)        # b(theta=1) == b
    assert np.allclose(tab.interp_weights(0.0), 0)


def _fixed_steps(f, u0, n):
    u, t, dt = u0, 0.0, 1.0 / n
    k1 = f(u, t)
    for _ in range(n):
        ks, u, _, _ = tsit5_stages(f, u, k1, t, dt)
        k1, t = ks[6], t + dt
    return u


def test_convergence_order_five():
    f = lambda u, t: -u + np.sin(3 * t)
    u0 = np.array([[1.0, 2.0]])
    ref = _fixed_steps(f, u0, 512)
    e1 = np.abs(_fixed_steps(f, u0, 8) - ref).max()
    e2 = np.abs(_fixed_steps(f, u0, 16) - ref).max()
    assert 4.5 < np.log2(e1 / e2) < 5.8


def test_interpolant_fourth_order():
    f = lambda u, t: -u
    u0 = np.array([[1.0]])
    errs = []
    for dt in (0.2, 0.1):
        ks, _, _, _ = tsit5_stages(f, u0, f(u0, 0.0), 0.0, dt)
        errs.append(abs(interp_step(u0, ks, dt, 0.37)[0, 0] - np.exp(-0.37 * dt)))
    assert 4.3 < np.log2(errs[0] / errs[1]) < 5.7


def test_fastpow_close_to_pow():
    for x in (1e-6, 3e-3, 0.05, 0.9, 1.0, 7.5, 300.0):
        for y in (0.14, 0.08):
            got = float(fastpow(np.float32(x), np.float32(y)))
            assert abs(got / x ** y - 1) < 2e-3
    assert fastpow(np.float32(0), np.float32(0.14)) == 0
    assert isinstance(fastpow(np.float32(0.3), np.float32(0.14)), np.float32)


def test_adaptive_solve_linear_system():
    A = np.array([[-0.5, 2.0], [-2.0, -0.5]])
    f = lambda u, t: A @ u
    u0 = np.array([[1.0], [0.0]])
    sol = solve_tsit5(f, u0, 0.0, 1.0, abstol=1e-9, reltol=1e-9, maxiters=10000)
    exact = np.exp(-0.5) * np.array([[np.cos(2.0)], [-np.sin(2.0)]])
    assert sol.retcode == 0 and sol.ts[-1] == 1.0
    assert np.abs(sol.us[-1] - exact).max() < 1e-8
    assert sol.nf == 3 + 6 * (sol.naccept + sol.nreject)
    assert np.abs(sol(0.4321) - np.exp(-0.5 * 0.4321) * np.array(
        [[np.cos(2 * 0.4321)], [-np.sin(2 * 0.4321)]])).max() < 1e-7
    # backwards in time retraces the trajectory
    back = solve_tsit5(f, sol.us[-1], 1.0, 0.0, abstol=1e-9, reltol=1e-9, maxiters=10000)
    assert back.tdir == -1 and np.abs(back.us[-1] - u0).max() < 1e-7


def test_tstops_and_retcodes():
    f = lambda u, t: -u
    u0 = np.ones((1, 1), np.float32)
    sol = solve_tsit5(f, u0, 0.0, 1.0, abstol=1e-6, reltol=1e-6, tstops=[0.5, 0.25])
    assert np.float32(0.25) in sol.ts and np.float32(0.5) in sol.ts
    sol = solve_tsit5(f, u0, 0.0, 1.0, abstol=1e-7, reltol=1e-7, maxiters=3)
    assert sol.retcode == 1 and sol.naccept + sol.nreject == 3
    bad = solve_tsit5(lambda u, t: u * np.float32(np.nan), u0, 0.0, 1.0, abstol=1e-6, reltol=1e-6)
    assert bad.retcode != 0


def _tiny(td=True, act="gelu", dtype=np.float64, seed=0, D=2, H=4, B=3):
    model = MLP([Dense(D, H, act), Dense(H, D, "identity")], time_dependent=td)
    rng = np.random.default_rng(seed)
    ps = glorot_uniform_params(model, rng, dtype)
    ps += (0.1 * rng.standard_normal(ps.size)).astype(dtype)      # non-zero biases
    x = rng.standard_normal((D, B)).astype(dtype)
    return model, ps, x


def test_vjp_matches_finite_differences():
    model, ps, x = _tiny()
    lam = np.random.default_rng(1).standard_normal(x.shape)
    a, dp = model.vjp(x, ps, 0.3, lam)
    eps = 1e-6
    for i in range(0, ps.size, 3):
        e = np.zeros_like(ps); e[i] = eps
        fd = ((model.f(x, ps + e, 0.3) - model.f(x, ps - e, 0.3)) * lam).sum() / (2 * eps)
        assert abs(fd - dp[i]) < 1e-7
    e = np.zeros_like(x); e[1, 2] = eps
    fd = ((model.f(x + e, ps, 0.3) - model.f(x - e, ps, 0.3)) * lam).sum() / (2 * eps)
    assert abs(fd - a[1, 2]) < 1e-7


@pytest.mark.parametrize("td", [True, False])
def test_adjoint_matches_finite_differences(td):
    model, ps, x = _tiny(td=td)
    tol = 1e-11
    c = np.random.default_rng(2).standard_normal(x.shape)

    def loss(ps_, x_):
        s = solve_tsit5(lambda u, t: model.f(u, ps_, t), x_, 0.0, 1.0, abstol=tol, reltol=tol,
                        maxiters=100000)
        return (c * s.us[-1]).sum(), s

    _, sol = loss(ps, x)
    d_x, d_ps, _ = adjoint_backward(model, ps, sol, [1.0], [c], abstol=tol, reltol=tol,
                                    maxiters=100000)
    eps = 1e-5
    for i in range(0, ps.size, 2):
        e = np.zeros_like(ps); e[i] = eps
        fd = (loss(ps + e, x)[0] - loss(ps - e, x)[0]) / (2 * eps)
        assert abs(fd - d_ps[i]) < 2e-6 * max(1, abs(fd))
    for idx in [(0, 0), (1, 2)]:
        e = np.zeros_like(x); e[idx] = eps
        fd = (loss(ps, x + e)[0] - loss(ps, x - e)[0]) / (2 * eps)
        assert abs(fd - d_x[idx]) < 2e-6 * max(1, abs(fd))


def test_adjoint_with_intermediate_cotangent():
    model, ps, x = _tiny()
    tol = 1e-11
    rng = np.random.default_rng(3)
    c1, c2 = rng.standard_normal(x.shape), rng.standard_normal(x.shape)
    tm = 0.4

    def loss(ps_):
        s = solve_tsit5(lambda u, t: model.f(u, ps_, t), x, 0.0, 1.0, abstol=tol, reltol=tol,
                        maxiters=100000)
        return (c1 * s(tm)).sum() + (c2 * s.us[-1]).sum(), s

    _, sol = loss(ps)
    _, d_ps, bsol = adjoint_backward(model, ps, sol, [tm, 1.0], [c1, c2], abstol=tol,
                                     reltol=tol, maxiters=100000)
    assert tm in bsol.ts
    eps = 1e-5
    for i in range(0, ps.size, 4):
        e = np.zeros_like(ps); e[i] = eps
        fd = (loss(ps + e)[0] - loss(ps - e)[0]) / (2 * eps)
        assert abs(fd - d_ps[i]) < 5e-6 * max(1, abs(fd))


@pytest.mark.parametrize("reg_type", ["error_estimate", "stiffness_estimate"])
def test_reg_step_backward_matches_finite_differences(reg_type):
    model, ps, x = _tiny(act="tanh")
    t1, dt, tol = 0.3, 0.17, 1e-3
    k1 = model.f(x, ps, t1)          # frozen: integrator creation is non-differentiable

    def reg(ps_):
        return perform_step_tsit5_reg(lambda u, t: model.f(u, ps_, t), x, k1, t1, dt, tol, tol,
                                      reg_type)[1]

    g = reg_step_backward(model, ps, x, k1, t1, dt, tol, tol, reg_type, 1.0)
    eps = 1e-6
    for i in range(0, ps.size, 2):
        e = np.zeros_like(ps); e[i] = eps
        fd = (reg(ps + e) - reg(ps - e)) / (2 * eps)
        assert abs(fd - g[i]) < 1e-5 * max(1e-3, abs(fd)), (i, fd, g[i])


# ---- the reference's own property tests (test/runtests.jl), ODE items ----------------
def _chain(mode, td):
    """Dense(2=>2,gelu) -> NeuralODE(...) -> diffeqsol_to_array -> Dense(2=>2)
    (runtests.jl:9-12)."""
    inner = MLP([Dense(2, 4, "gelu"), Dense(4, 2, "identity")], time_dependent=td)
    rng = np.random.default_rng(0)
    node = NeuralODE(inner, regularize=mode, tspan=(0.0, 1.0))
    ps_node = glorot_uniform_params(inner, rng)
    W1 = rng.uniform(-1, 1, (2, 2)).astype(np.float32)
    W2 = rng.uniform(-1, 1, (2, 2)).astype(np.float32)
    x = rng.standard_normal((2, 1)).astype(np.float32)
    st = node.initialstates(np.random.default_rng(0))
    return node, ps_node, W1, W2, x, st


@pytest.mark.parametrize("mode", ["none", "unbiased", "biased"])
@pytest.mark.parametrize("td", [True, False])
def test_reference_property_suite(mode, td):
    node, ps, W1, W2, x, st = _chain(mode, td)
    h = W1 @ x
    h = 0.5 * h * (1 + np.tanh(0.7978845608 * (h + 0.044715 * h ** 3)))
    sol, st2, aux = node.forward(h.astype(np.float32), ps, st)
    y = W2 @ diffeqsol_to_array(sol)
    assert y.dtype == np.float32 and y.shape == (2, 1)                 # :21
    if mode == "none":
        assert st2["reg_val"] == 0                                      # :22
    else:
        assert st2["reg_val"] != 0                                      # :118
    assert st2["nfe"] > 0 and st["nfe"] == -1
    # d sum(y) / d (x, ps)
    d_last = (W2.T @ np.ones((2, 1), np.float32))
    d_us = [None] * (len(sol.u) - 1) + [d_last]
    d_h, d_ps = node.backward(aux, d_us, 0.0, ps)
    assert np.all(np.isfinite(d_h)) and np.all(d_h != 0)               # :24-26
    assert np.all(np.isfinite(d_ps)) and np.all(d_ps != 0)             # :27-29
    if mode != "none":
        # d reg / d (x, ps): x-gradient is structurally zero ("=== nothing", :129)
        d_h2, d_ps2 = node.backward(aux, [None] * len(sol.u), 1.0, ps)
        assert np.all(d_h2 == 0)
        assert np.all(np.isfinite(d_ps2)) and np.any(d_ps2 != 0)       # :130-131


def test_layer_state_contract_and_nfe():
    node, ps, *_ , x, st = _chain("unbiased", True)
    sol, st2 = node(x, ps, st)
    assert set(st2) == {"model", "nfe", "reg_val", "rng", "training"}  # neural_ode.jl:83
    assert len(sol.u) == 2 and sol.t[1] == np.float32(1.0)             # [u(t1), u(t2)] :108
    _, _, aux = node.forward(x, ps, st)
    s = aux["sol"]
    assert st2["nfe"] == 3 + 6 * (s.naccept + s.nreject) + 9           # :79, perform_step.jl:31
    # eval mode falls back to the vanilla path (:66)
    st_eval = dict(st, training=False)
    sol_e, st_e = node(x, ps, st_eval)
    assert st_e["reg_val"] == 0 and len(sol_e.u) == 1 and st_e["nfe"] == s.nf
    # same rng state -> same t1; returned rng has advanced (:69,:83)
    _, st3 = node(x, ps, st)
    assert st3["reg_val"] == st2["reg_val"]
    _, st4 = node(x, ps, st2)
    assert st4["reg_val"] != st2["reg_val"]


def test_constructor_validation():
    inner = MLP([Dense(2, 4, "tanh"), Dense(4, 2)], time_dependent=False)
    with pytest.raises(ValueError):
        NeuralODE(inner, regularize="bogus")                           # utils.jl:53-58
    with pytest.raises(ValueError):
        NeuralODE(inner, regularize_type="bogus")
    assert NeuralODE(inner, regularize=True).regularize == "unbiased"  # neural_ode.jl:14-16
    assert NeuralODE(inner, regularize=False).regularize == "none"


def test_saveat_timeseries_and_correction():
    inner = MLP([Dense(3, 5, "tanh"), Dense(5, 3, "tanh")], time_dependent=False, input_act="tanh")
    rng = np.random.default_rng(0)
    ps = glorot_uniform_params(inner, rng)
    x = rng.standard_normal((3, 4)).astype(np.float32)
    saveat = [0.0, 0.2, 0.55, 1.0]
    node = NeuralODE(inner, regularize="unbiased", saveat=saveat, abstol=1e-6, reltol=1e-6)
    st = node.initialstates(np.random.default_rng(5))
    sol, st2, aux = node.forward(x, ps, st)
    assert [float(t) for t in sol.t] == [np.float32(s) for s in saveat]  # t1 column dropped
    ts = diffeqsol_to_timeseries(sol)
    assert ts.shape == (3, 4, 4) and np.array_equal(ts[:, 0, :], x)
    d_us = [np.ones_like(x) for _ in saveat]
    d_x, d_ps = node.backward(aux, d_us, 0.5, ps)
    assert np.all(np.isfinite(d_x)) and np.all(np.isfinite(d_ps))
    b = aux["bsol"]
    for s in (0.2, 0.55):
        assert np.float32(s) in b.ts                                   # lambda jumps are tstops


def test_initdt_matches_hand_computation():
    f = lambda u, t: -2 * u
    u0 = np.full((2, 2), 1.5, np.float32)
    dt, extra = ode_initdt(f, u0, 0.0, 1, 1.0, 1e-6, 1e-3)
    assert extra == 2 and 0 < dt <= 1
    sk = 1e-6 + 1.5 * 1e-3
    d0, d1 = 1.5 / sk, 3.0 / sk
    dt0 = 0.01 * d0 / d1
    d2 = (2 * (dt0 * 3.0)) / sk / dt0
    dt1 = 10 ** (-(2 + np.log10(max(d1, d2))) / 5)
    assert abs(dt - min(100 * dt0, dt1)) < 1e-5


def test_sosri_coefficients_and_deterministic_limit():
    c = SOSRI
    assert abs(c["alpha1"] + c["alpha2"] + c["alpha3"] + c["alpha4"] - 1) < 1e-12
    assert abs(sum(c[f"beta1{i}"] for i in range(1, 5)) - 1) < 1e-12
    for r in (2, 3, 4):
        assert abs(sum(c[f"beta{r}{i}"] for i in range(1, 5))) < 1e-12
    assert abs(c["a021"] - c["c02"]) < 1e-15 and abs(c["a031"] + c["a032"] - c["c03"]) < 1e-12
    # zero diffusion: one step of u' = -u is a consistent (order >= 2) RK step
    u0 = np.ones((1, 1))
    z = np.zeros((1, 1))
    errs = []
    for dt in (0.1, 0.05):
        u, reg, nfe, _ = perform_step_sosri_reg(lambda u, t: -u, lambda u, t: 0 * u, u0, 0.0, dt,
                                                z, z, 1e-3, 1e-3)
        errs.append(abs(u[0, 0] - np.exp(-dt)))
        assert nfe == 0 and reg > 0
    assert np.log2(errs[0] / errs[1]) > 2.5
