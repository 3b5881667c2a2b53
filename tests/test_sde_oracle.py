"""Self-validation of the SDE side of the CPU oracle (oracle/lrnde_sde_oracle.py).  CPU only.

Pinned pieces: Philox4x32-10 against the Random123 known-answer vectors; the in-tree SOSRI
step against its consistency conditions; everything un-vendored (loop, RSwM3, initdt) through
properties (Wiener statistics with rejections, gradients vs finite differences in Float64) and
the reference's own three NeuralDSDE property tests (test/runtests.jl:340-430)."""
import importlib.util
import os

import numpy as np
import pytest

import oracle as orc
from oracle import lrnde_sde_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_sde", os.path.join(HERE, "golden", "make_golden_sde.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def test_philox_known_answer_vectors():
    """Random123 kat_vectors, philox4x32 with 10 rounds."""
    def run(c, k):
        o = so.philox4x32_10(*[np.array([v], dtype=np.uint32) for v in c], k[0], k[1])
        return [int(v[0]) for v in o]
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normal_moments_and_streams():
    z = so.philox_normal(5, 0, 0, 200000, np.float64)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.01 and abs((z ** 4).mean() - 3) < 0.1
    z2 = so.philox_normal(5, 1, 0, 200000, np.float64)
    z3 = so.philox_normal(5, 0, 1, 200000, np.float64)
    assert abs(np.corrcoef(z, z2)[0, 1]) < 0.01 and abs(np.corrcoef(z, z3)[0, 1]) < 0.01
    assert np.array_equal(so.philox_normal(5, 0, 0, 100), so.philox_normal(5, 0, 0, 200000)[:100])   # counter based


def test_sosri_tableau_consistency():
    c = so.SOSRI
    assert abs(c["alpha1"] + c["alpha2"] + c["alpha3"] + c["alpha4"] - 1) < 1e-14
    assert abs(sum(c[f"beta1{i}"] for i in range(1, 5)) - 1) < 1e-13
    for b in (2, 3, 4):
        assert abs(sum(c[f"beta{b}{i}"] for i in range(1, 5))) < 1e-13
    assert abs(c["a021"] - c["c02"]) < 1e-15 and abs(c["a031"] + c["a032"] - c["c03"]) < 1e-14
    assert abs(c["a121"] - c["c12"]) < 1e-15 and abs(c["a141"] + c["a142"] + c["a143"] - c["c14"]) < 1e-14


def _nets(D=2, H=8, dact="tanh"):
    drift = orc.MLP([orc.Dense(D, H, "tanh"), orc.Dense(H, D, "identity")], time_dependent=False)
    diff = orc.MLP([orc.Dense(D, D, dact)], time_dependent=False)
    return drift, diff


def test_rswm3_keeps_the_wiener_process_consistent():
    """With rejections (qmax = 10) the accepted increments still sum to W(1) ~ N(0, 1) and the
    accepted dt's to the interval length."""
    drift, diff = _nets()
    rng = np.random.default_rng(0)
    pf = (3 * orc.glorot_uniform_params(drift, rng)).astype(np.float32)
    pg = (2 * orc.glorot_uniform_params(diff, rng) + 0.3).astype(np.float32)
    fd = lambda u, t: drift.f(u, pf, t)
    gd = lambda u, t: diff.f(u, pg, t)
    x = rng.standard_normal((2, 1)).astype(np.float32)
    tot, nrej = [], 0
    for seed in range(150):
        sol = so.solve_sosri(fd, gd, x, 0.0, 1.0, abstol=0.05, reltol=0.05, seed=seed, maxiters=10000,
                             consts=dict(qmax=10.0))
        assert sol.retcode == "Success" and sol.ts[-1] == np.float32(1.0)
        assert abs(sum(float(s[1]) for s in sol.steps) - 1.0) < 1e-5
        tot.append(sum(s[2] for s in sol.steps).ravel())
        tot.append(sum(s[3] for s in sol.steps).ravel())
        nrej += sum(1 for l in sol.log if not l[3])
    t = np.concatenate(tot)
    assert nrej > 20
    assert abs(t.var() - 1) < 0.2 and abs(t.mean()) < 0.15


def test_strong_convergence_to_a_fine_reference():
    """Geometric Brownian motion du = a u dt + b u dW: the adaptive solution approaches the exact
    exp((a - b^2/2) t + b W(t)) as the tolerance tightens."""
    a, b = 0.4, 0.5
    fd = lambda u, t: np.float64(a) * u
    gd = lambda u, t: np.float64(b) * u
    x = np.ones((1, 64))
    errs = []
    for tol in (3e-2, 3e-3):
        sol = so.solve_sosri(fd, gd, x, 0.0, 1.0, abstol=tol, reltol=tol, seed=11, maxiters=100000, pow_mode="exact")
        W = sum(s[2] for s in sol.steps)
        exact = np.exp((a - b * b / 2) + b * W)
        errs.append(np.abs(sol.us[-1] - exact).mean())
    assert errs[1] < errs[0] and errs[1] < 5e-3


def test_backward_matches_finite_differences_float64():
    D, B = 5, 4
    drift, diff = _nets(D, 7)
    rng = np.random.default_rng(1)
    ps = np.concatenate([orc.glorot_uniform_params(drift, rng, np.float64),
                         0.4 * orc.glorot_uniform_params(diff, rng, np.float64)])
    ps += 0.05 * rng.standard_normal(ps.size)
    x = rng.standard_normal((D, B))
    n = so.NeuralDSDE(drift, diff, regularize="unbiased", abstol=0.05, reltol=0.05, dtype=np.float64, seed=3,
                      pow_mode="exact", saveat=[0.3, 1.0])
    out, st2, aux = n.forward(x, ps, n.initialstates(np.random.default_rng(1)))
    sol = aux["sol"]
    cots = [rng.standard_normal((D, B)) for _ in out.u]
    dreg = 0.7
    dx, dps = n.backward(aux, cots, dreg, ps)

    def loss(ps_, x_, with_reg=True):
        pf, pg = n.split(ps_)
        fd = lambda u, t: drift.f(u, pf, t)
        gd = lambda u, t: diff.f(u, pg, t)
        us = [x_]
        for (t, dt, dW, dZ) in sol.steps:          # frozen step sequence and noise (TrackerAdjoint)
            u, _ = so.sosri_step(fd, gd, us[-1], t, dt, dW, dZ, n.abstol, n.reltol, so.SDE_CONSTS["delta"])
            us.append(u)
        s2 = so.SDESolution(ts=sol.ts, us=us)
        L = sum(np.sum(c * s2(s)) for c, s in zip(cots, out.t))
        if with_reg:
            _, reg, _, _ = orc.perform_step_sosri_reg(fd, gd, aux["u1"], aux["t1"], aux["dt_reg"], aux["dW_reg"],
                                                      aux["dZ_reg"], n.abstol, n.reltol, so.SDE_CONSTS["delta"])
            L = L + dreg * reg
        return L
    eps = 1e-6
    for i in rng.choice(ps.size, 10, replace=False):
        p1, p2 = ps.copy(), ps.copy()
        p1[i] += eps; p2[i] -= eps
        fdv = (loss(p1, x) - loss(p2, x)) / (2 * eps)
        assert abs(dps[i] - fdv) < 1e-6 * max(1.0, abs(fdv))
    for (a, b) in [(0, 0), (3, 2), (4, 3)]:
        x1, x2 = x.copy(), x.copy()
        x1[a, b] += eps; x2[a, b] -= eps
        fdv = (loss(ps, x1) - loss(ps, x2)) / (2 * eps)      # the regulariser does not depend on x
        assert abs(dx[a, b] - fdv) < 1e-6 * max(1.0, abs(fdv))


@pytest.mark.parametrize("mode", ["none", "unbiased", "biased"])
def test_reference_property_tests_neuraldsde(mode):
    """test/runtests.jl:340-430: Dense(2=>2, gelu) -> NeuralDSDE(Chain(Dense(2=>4, gelu), Dense(4=>2)),
    Dense(2=>2)) -> diffeqsol_to_array -> Dense(2=>2), B = 1."""
    rng = np.random.default_rng(0)
    drift = orc.MLP([orc.Dense(2, 4, "gelu"), orc.Dense(4, 2, "identity")], time_dependent=False)
    diff = orc.MLP([orc.Dense(2, 2, "identity")], time_dependent=False)
    n = so.NeuralDSDE(drift, diff, regularize=mode, tspan=(0.0, 1.0), seed=0)
    ps = np.concatenate([orc.glorot_uniform_params(drift, rng), orc.glorot_uniform_params(diff, rng)])
    W1 = rng.standard_normal((2, 2)).astype(np.float32); W2 = rng.standard_normal((2, 2)).astype(np.float32)
    x = rng.standard_normal((2, 1)).astype(np.float32)
    h = orc.lrnde_oracle._act("gelu", W1 @ x)
    out, st2, aux = n.forward(h, ps, n.initialstates(np.random.default_rng(0)))
    y = W2 @ orc.diffeqsol_to_array(out)
    assert y.dtype == np.float32 and y.shape == (2, 1)
    assert (st2["reg_val"] == 0) == (mode == "none")
    assert st2["nfe_drift"] > 0 and st2["nfe_diffusion"] > 0
    d_us = [None] * (len(out.u) - 1) + [W2.T @ np.ones((2, 1), np.float32)]
    dh, dps = n.backward(aux, d_us, 0.0, ps)
    assert np.all(np.isfinite(dh)) and np.all(dh != 0)
    assert np.all(np.isfinite(dps)) and np.all(dps != 0)
    if mode != "none":
        dh2, dps2 = n.backward(aux, [None] * len(out.u), 1.0, ps)
        assert np.all(dh2 == 0)                         # gs_x === nothing (runtests.jl:394,426)
        assert np.all(np.isfinite(dps2)) and np.any(dps2 != 0)


def test_invalid_regularize_raises():
    drift, diff = _nets()
    with pytest.raises(ValueError):
        so.NeuralDSDE(drift, diff, regularize="sometimes")          # utils.jl:53-58


def test_other_in_tree_steps_reduce_to_their_deterministic_limits():
    """perform_step.jl:108-206 with zero diffusion: RKMilCommute -> explicit Euler, LambaEulerHeun ->
    Heun; with dW = 0 and non-zero diffusion both keep u finite and return EEst * dt."""
    fd = lambda u, t: -u + np.float32(0.3)
    g0 = lambda u, t: np.zeros_like(u)
    u0 = np.array([[1.0, -2.0, 0.5]], np.float32).T
    dt = np.float32(0.1)
    z = np.zeros_like(u0)
    u, reg, _, _ = so.perform_step_rkmil_reg(fd, g0, u0, 0.0, dt, z, z, 1e-2, 1e-2)
    assert np.allclose(u, u0 + dt * fd(u0, 0.0), atol=1e-7) and reg > 0
    u, reg, _, _ = so.perform_step_lamba_eulerheun_reg(fd, g0, u0, 0.0, dt, z, z, 1e-2, 1e-2)
    k1 = fd(u0, 0.0); k2 = fd(u0 + dt * k1, dt)
    assert np.allclose(u, u0 + dt / 2 * (k1 + k2), atol=1e-7) and reg > 0
    gd = lambda u, t: np.float32(0.5) * u
    dW = np.full_like(u0, 0.2)
    for step in (so.perform_step_rkmil_reg, so.perform_step_lamba_eulerheun_reg):
        u, reg, nf, dto = step(fd, gd, u0, 0.0, dt, dW, z, 1e-2, 1e-2)
        assert np.all(np.isfinite(u)) and np.isfinite(reg) and nf == 0 and dto == dt
    # Milstein correction: u = K + L dW + (g(K + sqrt(dt) L) - L)/sqrt(dt) * (dW^2 - dt)/2 ~ exact for g = b u
    u, _, _, _ = so.perform_step_rkmil_reg(lambda u, t: np.zeros_like(u), gd, u0, 0.0, dt, dW, z, 1e-2, 1e-2)
    assert np.allclose(u, u0 * (1 + 0.5 * 0.2 + 0.5 * 0.25 * (0.04 - 0.1)), atol=1e-6)


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_golden_sde_fixtures_reproduce(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    r = mg.run(name)
    assert np.array_equal(r["log"][:, 3], g["log"][:, 3])
    assert np.allclose(r["log"][:, :3], g["log"][:, :3], rtol=1e-5)
    assert np.array_equal(r["nfe"], g["nfe"]) and int(r["ndraws"]) == int(g["ndraws"])
    assert np.allclose(r["u"], g["u"], rtol=1e-5, atol=1e-6)
    assert np.allclose(r["dps"], g["dps"], rtol=1e-4, atol=1e-6)
    assert np.isclose(r["reg"], g["reg"], rtol=1e-4)
