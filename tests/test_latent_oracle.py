"""Self-validation of the latent-ODE encoder oracle (oracle/lrnde_latent_oracle.py).  CPU only."""
import numpy as np

import oracle as orc


def _inputs(F=9, H=6, L=4, T=5, B=3, dtype=np.float64, seed=0):
    rng = np.random.default_rng(seed)
    ps = orc.gru_init(rng, F, H, L, dtype) + 0.1 * rng.standard_normal(orc.gru_nparams(F, H, L)).astype(dtype)
    x = rng.standard_normal((F, T, B)).astype(dtype)
    x[F // 2:, :, :] = (rng.random((F - F // 2, T, B)) < 0.3)
    x[F // 2:, 2, 1] = 0           # one (time, sample) with nothing observed and dt = 0
    return ps, x, rng


def test_gru_backward_matches_finite_differences():
    F, H, L = 9, 6, 4
    ps, x, rng = _inputs(F, H, L)
    y, car = orc.gru_recurrence(ps, x, F, H, L)
    dy = rng.standard_normal(y.shape)
    dps = orc.gru_recurrence_backward(ps, x, F, H, L, car, dy)
    eps = 1e-6
    for i in rng.choice(ps.size, 16, replace=False):
        p1, p2 = ps.copy(), ps.copy()
        p1[i] += eps; p2[i] -= eps
        fd = (np.sum(dy * orc.gru_recurrence(p1, x, F, H, L)[0]) - np.sum(dy * orc.gru_recurrence(p2, x, F, H, L)[0])) / (2 * eps)
        assert abs(dps[i] - fd) < 1e-7 * max(1.0, abs(fd))


def test_gru_cell_quirks_of_the_reference():
    """latent_ode.jl:37: new_y_mean uses new_state_std, so the first half of the new-state network's output
    never reaches the result; :40-43: an unobserved time point (mask rows and dt all zero) keeps the carry."""
    F, H, L = 9, 6, 4
    ps, x, _ = _inputs(F, H, L)
    P = orc.GRUParams(ps, F, H, L)
    W2 = P.nets[2][1][0]
    y0, _ = orc.gru_recurrence(ps, x, F, H, L)
    W2[:L, :] += 10.0                              # perturb the rows that produce new_state_mean (a view into ps)
    y1, _ = orc.gru_recurrence(ps, x, F, H, L)
    assert np.array_equal(y0, y1)
    ym, ys = np.full((L, 3), 0.3), np.full((L, 3), 0.7)
    nm, nsd = orc.latent_gru_cell(P, x[:, 2, :], ym, ys)
    assert np.array_equal(nm[:, 1], ym[:, 1]) and np.array_equal(nsd[:, 1], ys[:, 1])
    assert not np.array_equal(nm[:, 0], ym[:, 0])
    # the first call starts from y_mean = 0, y_std = 1 (:20-24) and returns vcat(mean, std)
    y, car = orc.gru_recurrence(ps, x, F, H, L)
    assert y.shape == (2 * L, 3) and np.all(car[0][0] == 0) and np.all(car[0][1] == 1)


def test_reparameterize_and_losses():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((8, 5)).astype(np.float32)
    eps = rng.standard_normal((4, 5)).astype(np.float32)
    y, mu, ls = orc.reparameterize(x, eps, True)
    assert np.allclose(y, x[:4] + np.exp(x[4:] / 2) * eps) and mu is not None and np.array_equal(ls, x[4:])
    y2, mu2, ls2 = orc.reparameterize(x, eps, False)            # common.jl:73-77: eval mode returns mu three times
    assert np.array_equal(y2, x[:4]) and np.array_equal(ls2, x[:4])
    assert np.allclose(orc.kl_divergence(np.zeros((4, 5), np.float32), np.zeros((4, 5), np.float32)), 0)
    d = np.zeros((3, 6, 5), np.float32)
    m = np.ones_like(d)
    ll = orc.log_likelihood_loss(d, m)                          # zero residual: the Gaussian normaliser only
    assert np.allclose(ll, -np.log(0.01) - np.log(2 * np.pi) / 2, rtol=1e-6)
