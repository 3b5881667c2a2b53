"""CPU oracle, latent-ODE encoder side.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy Float32 (Float64 twin by dtype), the in-tree code either side of the ODE solve of
the physionet config:

* ``LatentGRUCell``          <- src/layers/latent_ode.jl:1-48 (quirks kept: ``new_y_mean`` is built from
                                ``new_state_std`` (:37), and the observation mask sums the mask rows AND the
                                trailing dt row, ``x[(size(x,1) div 2 + 1):end, :]`` (:40))
* ``Recurrence(cell)``       <- Lux.Recurrence as used at experiments/src/construct.jl:231 (the cell is
                                applied along the time dimension ``ndims - 1``; the last output is returned)
* ``ReparameterizeLayer``    <- src/layers/common.jl:48-77
* ``log_likelihood_loss`` / ``kl_divergence`` <- experiments/src/utils.jl:94-101

The reverse pass is hand-derived and checked against central differences in Float64
(tests/test_latent_oracle.py)."""
from __future__ import annotations

import numpy as np

from .lrnde_oracle import _act, _act_grad

__all__ = ["GRUParams", "latent_gru_cell", "gru_recurrence", "gru_recurrence_backward", "reparameterize",
           "log_likelihood_loss", "kl_divergence", "gru_nparams", "gru_init"]


def gru_nparams(F, H, L):
    I = 2 * L + F
    return 2 * ((I + 1) * H + (H + 1) * L) + (I + 1) * H + (H + 1) * 2 * L


class GRUParams:
    """Flat ComponentArray order: update_gate, reset_gate, new_state; each a Chain(Dense, Dense) with
    ``layer_k.weight[out x in]`` (column-major) followed by ``layer_k.bias``."""

    def __init__(self, ps, F, H, L):
        I = 2 * L + F
        self.F, self.H, self.L, self.I = F, H, L, I
        off = 0
        self.nets = []
        self.offs = []
        for out2, act2 in ((L, "sigmoid"), (L, "sigmoid"), (2 * L, "tanh")):
            layers = []
            for (i, o, a) in ((I, H, "tanh"), (H, out2, act2)):
                W = ps[off:off + o * i].reshape((o, i), order="F"); wo = off; off += o * i
                b = ps[off:off + o]; bo = off; off += o
                layers.append((W, b, a, wo, bo))
            self.nets.append(layers)
        assert off == ps.size, (off, ps.size)


def _net_fwd(layers, x, cache=None):
    for (W, b, a, _, _) in layers:
        pre = W @ x + b[:, None]
        y = _act(a, pre)
        if cache is not None:
            cache.append((x, pre, y))
        x = y
    return x


def _net_vjp(layers, x, cot, dps):
    cache = []
    _net_fwd(layers, x, cache)
    g = cot
    for (W, b, a, wo, bo), (xin, pre, y) in reversed(list(zip(layers, cache))):
        d = g * _act_grad(a, pre, y)
        dps[wo:wo + W.size] += (d @ xin.T).ravel(order="F")
        dps[bo:bo + b.size] += d.sum(axis=1)
        g = W.T @ d
    return g


def gru_init(rng, F, H, L, dtype=np.float32):
    """Lux defaults: Glorot-uniform weights, zero biases."""
    I = 2 * L + F
    out = []
    for out2 in (L, L, 2 * L):
        for (i, o) in ((I, H), (H, out2)):
            a = np.sqrt(6.0 / (i + o))
            out.append(rng.uniform(-a, a, size=(o, i)).astype(dtype).ravel(order="F"))
            out.append(np.zeros(o, dtype))
    return np.concatenate(out)


def latent_gru_cell(P: GRUParams, x, y_mean, y_std):
    """latent_ode.jl:26-48.  Returns (new_y_mean, new_y_std)."""
    T = x.dtype.type
    L = P.L
    yc = np.concatenate([y_mean, y_std, x], axis=0)
    ug = _net_fwd(P.nets[0], yc)
    rg = _net_fwd(P.nets[1], yc)
    c2 = np.concatenate([y_mean * rg, y_std * rg, x], axis=0)
    ns = _net_fwd(P.nets[2], c2)
    nss = ns[L:, :]
    nm = (T(1) - ug) * nss + ug * y_mean          # :37 -- new_state_std, not new_state_mean
    nsd = (T(1) - ug) * nss + ug * y_std
    mask = (x[(x.shape[0] // 2):, :].sum(axis=0, keepdims=True) > 0).astype(x.dtype)   # :40
    return mask * nm + (T(1) - mask) * y_mean, mask * nsd + (T(1) - mask) * y_std


def gru_recurrence(ps, x, F, H, L):
    """Recurrence(LatentGRUCell(...)) over x [F, T, B]; returns (y [2L, B], carries) where
    carries[t] = (y_mean, y_std) BEFORE step t (what the reverse pass needs)."""
    P = GRUParams(ps, F, H, L)
    B = x.shape[2]
    ym = np.zeros((L, B), x.dtype)                # latent_ode.jl:20-24
    ys = np.ones((L, B), x.dtype)
    carries = []
    for t in range(x.shape[1]):
        carries.append((ym, ys))
        ym, ys = latent_gru_cell(P, x[:, t, :], ym, ys)
    return np.concatenate([ym, ys], axis=0), carries


def gru_recurrence_backward(ps, x, F, H, L, carries, d_y):
    """Pullback of gru_recurrence w.r.t. the parameters (the inputs are data)."""
    P = GRUParams(ps, F, H, L)
    T = x.dtype.type
    dps = np.zeros_like(ps)
    mb, sb = d_y[:L, :].copy(), d_y[L:, :].copy()
    for t in range(x.shape[1] - 1, -1, -1):
        ym, ys = carries[t]
        xt = x[:, t, :]
        yc = np.concatenate([ym, ys, xt], axis=0)
        ug = _net_fwd(P.nets[0], yc)
        rg = _net_fwd(P.nets[1], yc)
        c2 = np.concatenate([ym * rg, ys * rg, xt], axis=0)
        ns = _net_fwd(P.nets[2], c2)
        nss = ns[L:, :]
        mask = (xt[(xt.shape[0] // 2):, :].sum(axis=0, keepdims=True) > 0).astype(x.dtype)
        me, se = mask * mb, mask * sb
        nssb = (T(1) - ug) * (me + se)
        ugb = (ym - nss) * me + (ys - nss) * se
        pm = (T(1) - mask) * mb + ug * me
        psd = (T(1) - mask) * sb + ug * se
        nsb = np.concatenate([np.zeros_like(nssb), nssb], axis=0)
        c2b = _net_vjp(P.nets[2], c2, nsb, dps)
        pm = pm + rg * c2b[:L]
        psd = psd + rg * c2b[L:2 * L]
        rgb = ym * c2b[:L] + ys * c2b[L:2 * L]
        ycb = _net_vjp(P.nets[1], yc, rgb, dps) + _net_vjp(P.nets[0], yc, ugb, dps)
        mb, sb = pm + ycb[:L], psd + ycb[L:2 * L]
    return dps


def reparameterize(x, eps, training=True):
    """common.jl:57-77: (y, mu0, logsigma2).  ``eps`` replaces randn_like(rng, ...)."""
    L = x.shape[0] // 2
    mu = x[:L, :]
    if not training:
        return mu, mu, mu
    ls = x[L:, :]
    return mu + np.exp(ls / x.dtype.type(2)) * eps, mu, ls


def log_likelihood_loss(dpred, mask):
    """experiments/src/utils.jl:94-98; arrays [F, T, B] -> per-sample vector."""
    T = dpred.dtype.type
    s = T(0.01)
    lik = -(dpred ** 2) / (T(2) * s * s) - np.log(s) - np.log(T(2 * np.pi)) / T(2)
    return lik.sum(axis=(0, 1)) / mask.sum(axis=(0, 1))


def kl_divergence(mu, logsigma2):
    """experiments/src/utils.jl:101 (standard Gaussian prior)."""
    T = mu.dtype.type
    return (np.exp(logsigma2) + mu ** 2 - T(1) - logsigma2).mean(axis=0) / T(2)
