"""CPU restatement of the conv dynamics of the cifar10 config -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product path never does.  PARITY UNPINNED (see lrnde_oracle.py): the layer arithmetic
below lives in un-vendored Lux / NNlib and is restated from their published semantics.

Reference (file:line under /root/reference):
  * model      experiments/src/construct.jl:212-218 -- TDChain(Chain(Chain(Conv((3,3), 9=>64; pad=1,
               use_bias=false), BatchNorm(64, gelu)), Chain(Conv 65=>64, BatchNorm(64, gelu)), Conv 65=>8))
  * TDChain    src/layers/common.jl:10-45 -- ``t .* ones`` concatenated on dim ndims-1 (the channel dimension
               of a WHCN array) before EVERY outer layer, so each Conv sees in_ch + 1 channels
  * closure    src/layers/neural_ode.jl:44-48 (``dudt``); its pullback is what ZygoteVJP calls (:11)
  * Conv       NNlib ``conv``: TRUE convolution (flipped kernel), weight [kx,ky,cin,cout] column-major,
               zero padding 1:  y[o] = sum_k w[k] x[o + 1 - k]     [RECALLED, SURVEY A.7]
  * BatchNorm  training mode: batch mean / biased variance over (W,H,B), eps = 1f-5, affine (scale, bias),
               then the activation                                  [RECALLED, SURVEY A.7]

The state is the flat column-major [D, B] array the solver works on, D = W*H*C with w fastest.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .lrnde_oracle import _act, _act_grad

__all__ = ["ConvLayer", "ConvNet", "glorot_uniform_conv_params", "initial_conv_state", "cifar10_node_core",
           "conv2d", "conv2d_vjp", "batchnorm", "batchnorm_vjp", "augmenter", "augmenter_vjp"]


@dataclass(frozen=True)
class ConvLayer:
    """``Conv((3,3), in_ch(+1) => out_ch; pad=1, use_bias=false)`` [+ ``BatchNorm(out_ch, act)``]."""
    in_ch: int
    out_ch: int
    batchnorm: bool = False
    act: str = "identity"


@dataclass
class ConvNet:
    layers: List[ConvLayer]
    width: int
    height: int
    time_dependent: bool = True
    eps: float = 1e-5
    momentum: float = 0.1
    # st.model of the layer (Lux BatchNorm states): per BatchNorm layer running_mean[C] then running_var[C].
    # ``running`` set and ``track``: every training-mode f call updates it, as the ``dudt`` closure does through
    # its captured ``st_`` (neural_ode.jl:44-47); ``testmode``: normalise with it (Lux.testmode).
    running: Optional[np.ndarray] = None
    track: bool = False
    testmode: bool = False

    def __post_init__(self):
        td = 1 if self.time_dependent else 0
        off = 0
        self.offsets = []
        for L in self.layers:
            w_off = off
            off += 9 * (L.in_ch + td) * L.out_ch
            g_off = off
            if L.batchnorm:
                off += 2 * L.out_ch
            self.offsets.append((w_off, g_off))
        self.nparams = off
        soff = 0
        self.state_offsets = []
        for L in self.layers:
            self.state_offsets.append(soff)
            if L.batchnorm:
                soff += 2 * L.out_ch
        self.nstate = soff
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            assert a.out_ch == b.in_ch
        assert self.layers[0].in_ch == self.layers[-1].out_ch
        assert not self.layers[-1].batchnorm and self.layers[-1].act == "identity"

    @property
    def state_dims(self) -> int:
        return self.width * self.height * self.layers[0].in_ch

    # ---- layout helpers: flat [D, B] (w fastest) <-> [B, C, H, W]
    def _to_img(self, u, C):
        return np.ascontiguousarray(u.T).reshape(u.shape[1], C, self.height, self.width)

    def _to_flat(self, x):
        return np.ascontiguousarray(x.reshape(x.shape[0], -1).T)

    def unpack(self, ps):
        td = 1 if self.time_dependent else 0
        out = []
        for L, (wo, go) in zip(self.layers, self.offsets):
            n = 9 * (L.in_ch + td) * L.out_ch
            W = ps[wo:wo + n].reshape((3, 3, L.in_ch + td, L.out_ch), order="F")
            gamma = ps[go:go + L.out_ch] if L.batchnorm else None
            beta = ps[go + L.out_ch:go + 2 * L.out_ch] if L.batchnorm else None
            out.append((W, gamma, beta))
        return out

    @staticmethod
    def _conv(x, W):
        """x [B,Ci,H,W], W [kx,ky,ci,co] -> [B,Co,H,W]: true convolution, zero padding 1."""
        B, Ci, H, Wd = x.shape
        xp = np.zeros((B, Ci, H + 2, Wd + 2), x.dtype)
        xp[:, :, 1:-1, 1:-1] = x
        y = np.zeros((B, W.shape[3], H, Wd), x.dtype)
        for kx in range(3):
            for ky in range(3):
                y += np.einsum("io,bihw->bohw", W[kx, ky], xp[:, :, 2 - ky:2 - ky + H, 2 - kx:2 - kx + Wd],
                               optimize=True).astype(x.dtype)
        return y

    @staticmethod
    def _conv_vjp(x, W, d):
        """pullback of _conv: (g_x, dW)."""
        B, Ci, H, Wd = x.shape
        xp = np.zeros((B, Ci, H + 2, Wd + 2), x.dtype)
        xp[:, :, 1:-1, 1:-1] = x
        gp = np.zeros_like(xp)
        dW = np.zeros_like(W)
        for kx in range(3):
            for ky in range(3):
                sl = (slice(None), slice(None), slice(2 - ky, 2 - ky + H), slice(2 - kx, 2 - kx + Wd))
                dW[kx, ky] = np.einsum("bihw,bohw->io", xp[sl], d, optimize=True)
                gp[sl] += np.einsum("io,bohw->bihw", W[kx, ky], d, optimize=True).astype(x.dtype)
        return gp[:, :, 1:-1, 1:-1], dW

    def _with_time(self, x, t):
        if not self.time_dependent:
            return x
        tc = np.full((x.shape[0], 1) + x.shape[2:], x.dtype.type(t), dtype=x.dtype)
        return np.concatenate([x, tc], axis=1)          # common.jl:19-33: time channel appended last

    # f(u, p, t): the ``dudt`` closure of neural_ode.jl:45-48
    def f(self, u, ps, t, cache: Optional[list] = None):
        T = u.dtype.type
        x = self._to_img(u, self.layers[0].in_ch)
        for L, so, (W, gamma, beta) in zip(self.layers, self.state_offsets, self.unpack(ps)):
            xin = self._with_time(x, t)
            z = self._conv(xin, W)
            if L.batchnorm:
                C = L.out_ch
                if self.testmode:
                    mu = self.running[so:so + C].astype(z.dtype)[None, :, None, None]
                    var = self.running[so + C:so + 2 * C].astype(z.dtype)[None, :, None, None]
                else:
                    mu = z.mean(axis=(0, 2, 3), keepdims=True, dtype=z.dtype)
                    var = ((z - mu) ** 2).mean(axis=(0, 2, 3), keepdims=True, dtype=z.dtype)
                    if self.running is not None and self.track and cache is None:
                        n = z.size // C                   # Lux: running_var tracks the unbiased batch variance
                        m = self.running.dtype.type(self.momentum)
                        self.running[so:so + C] = (1 - m) * self.running[so:so + C] + m * mu.ravel()
                        self.running[so + C:so + 2 * C] = ((1 - m) * self.running[so + C:so + 2 * C]
                                                           + m * var.ravel() * (n / (n - 1)))
                invstd = T(1) / np.sqrt(var + T(self.eps))
                xhat = (z - mu) * invstd
                pre = gamma[None, :, None, None] * xhat + beta[None, :, None, None]
            else:
                invstd = xhat = None
                pre = z
            y = _act(L.act, pre)
            if cache is not None:
                cache.append((xin, xhat, invstd, pre, y))
            x = y
        return self._to_flat(x)

    def vjp(self, u, ps, t, lam):
        """(J_u^T lam, J_p^T lam) of f at (u, t)."""
        cache: list = []
        self.f(u, ps, t, cache)
        dps = np.zeros(self.nparams, dtype=u.dtype)
        g = self._to_img(lam, self.layers[-1].out_ch)
        for L, (wo, go), (W, gamma, _beta), (xin, xhat, invstd, pre, y) in reversed(list(zip(
                self.layers, self.offsets, self.unpack(ps), cache))):
            d = g * _act_grad(L.act, pre, y)
            if L.batchnorm:
                dps[go:go + L.out_ch] = (d * xhat).sum(axis=(0, 2, 3))
                dps[go + L.out_ch:go + 2 * L.out_ch] = d.sum(axis=(0, 2, 3))
                if self.testmode:                       # statistics are constants
                    d = gamma[None, :, None, None] * invstd * d
                else:
                    m1 = d.mean(axis=(0, 2, 3), keepdims=True, dtype=d.dtype)
                    m2 = (d * xhat).mean(axis=(0, 2, 3), keepdims=True, dtype=d.dtype)
                    d = gamma[None, :, None, None] * invstd * (d - m1 - xhat * m2)
            gx, dW = self._conv_vjp(xin, W, d)
            dps[wo:wo + dW.size] = dW.ravel(order="F")
            g = gx[:, :L.in_ch] if self.time_dependent else gx
        return self._to_flat(g), dps


def initial_conv_state(model: ConvNet, dtype=np.float32):
    """Lux.initialstates of the BatchNorm layers: running_mean = 0, running_var = 1."""
    st = np.zeros(model.nstate, dtype)
    for L, so in zip(model.layers, model.state_offsets):
        if L.batchnorm:
            st[so + L.out_ch:so + 2 * L.out_ch] = 1
    return st


def glorot_uniform_conv_params(model: ConvNet, rng: np.random.Generator, dtype=np.float32, jitter: float = 0.0):
    """Lux defaults: Conv weight Glorot-uniform (fan_in = 9*cin, fan_out = 9*cout), BatchNorm scale = 1, bias = 0.
    ``jitter`` perturbs scale / bias so that parity tests exercise their gradients away from the default point."""
    td = 1 if model.time_dependent else 0
    ps = np.zeros(model.nparams, dtype=dtype)
    for L, (wo, go) in zip(model.layers, model.offsets):
        cin = L.in_ch + td
        a = math.sqrt(6.0 / (9 * cin + 9 * L.out_ch))
        n = 9 * cin * L.out_ch
        ps[wo:wo + n] = rng.uniform(-a, a, size=n).astype(dtype)
        if L.batchnorm:
            ps[go:go + L.out_ch] = 1.0 + jitter * rng.standard_normal(L.out_ch)
            ps[go + L.out_ch:go + 2 * L.out_ch] = jitter * rng.standard_normal(L.out_ch)
    return ps


def cifar10_node_core(width=32, height=32) -> ConvNet:
    """experiments/src/construct.jl:212-218 (state 32x32x8 after the AugmenterLayer 3 -> 3+5 channels)."""
    return ConvNet([ConvLayer(8, 64, True, "gelu"), ConvLayer(64, 64, True, "gelu"), ConvLayer(64, 8)],
                   width, height, time_dependent=True)


# --------------------------------------------------------------------------
# Layers either side of the cifar10 NeuralODE (experiments/src/construct.jl:220-227), on WHCN arrays (W, H, C, B)
# --------------------------------------------------------------------------
def _split_conv_ps(ps, cin, cout, use_bias):
    n = 9 * cin * cout
    W = ps[:n].reshape((3, 3, cin, cout), order="F")
    return W, (ps[n:n + cout] if use_bias else None)


def conv2d(x, ps, out_ch, act="identity", use_bias=True):
    """Lux ``Conv((3,3), in => out, act; pad=(1,1))``: ``act.(conv(x, W) .+ b)``."""
    xb = np.ascontiguousarray(x.transpose(3, 2, 1, 0))
    W, b = _split_conv_ps(ps, xb.shape[1], out_ch, use_bias)
    z = ConvNet._conv(xb, W)
    if use_bias:
        z = z + b[None, :, None, None]
    return _act(act, z).transpose(3, 2, 1, 0)


def conv2d_vjp(x, ps, d_y, act="identity", use_bias=True):
    xb = np.ascontiguousarray(x.transpose(3, 2, 1, 0))
    dyb = np.ascontiguousarray(d_y.transpose(3, 2, 1, 0))
    cout = dyb.shape[1]
    W, b = _split_conv_ps(ps, xb.shape[1], cout, use_bias)
    z = ConvNet._conv(xb, W)
    if use_bias:
        z = z + b[None, :, None, None]
    d = dyb * _act_grad(act, z, _act(act, z))
    gx, dW = ConvNet._conv_vjp(xb, W, d)
    d_ps = np.concatenate([dW.ravel(order="F")] + ([d.sum(axis=(0, 2, 3))] if use_bias else []))
    return gx.transpose(3, 2, 1, 0), d_ps.astype(x.dtype)


def batchnorm(x, ps, act="identity", running=None, training=True, eps=1e-5, momentum=0.1):
    """Lux ``BatchNorm(C, act)``: returns ``(y, running')``; ``running`` = [mean; var]."""
    T = x.dtype.type
    C = x.shape[2]
    gamma, beta = ps[:C][None, None, :, None], ps[C:2 * C][None, None, :, None]
    if training:
        mu = x.mean(axis=(0, 1, 3), keepdims=True, dtype=x.dtype)
        var = ((x - mu) ** 2).mean(axis=(0, 1, 3), keepdims=True, dtype=x.dtype)
        if running is not None:
            n = x.size // C
            m = running.dtype.type(momentum)
            running = np.concatenate([(1 - m) * running[:C] + m * mu.ravel(),
                                      (1 - m) * running[C:] + m * var.ravel() * (n / (n - 1))]).astype(running.dtype)
    else:
        mu = running[:C].astype(x.dtype)[None, None, :, None]
        var = running[C:].astype(x.dtype)[None, None, :, None]
    xhat = (x - mu) / np.sqrt(var + T(eps))
    return _act(act, gamma * xhat + beta), running


def batchnorm_vjp(x, ps, d_y, act="identity", running=None, training=True, eps=1e-5):
    T = x.dtype.type
    C = x.shape[2]
    gamma, beta = ps[:C][None, None, :, None], ps[C:2 * C][None, None, :, None]
    if training:
        mu = x.mean(axis=(0, 1, 3), keepdims=True, dtype=x.dtype)
        var = ((x - mu) ** 2).mean(axis=(0, 1, 3), keepdims=True, dtype=x.dtype)
    else:
        mu = running[:C].astype(x.dtype)[None, None, :, None]
        var = running[C:].astype(x.dtype)[None, None, :, None]
    invstd = T(1) / np.sqrt(var + T(eps))
    xhat = (x - mu) * invstd
    pre = gamma * xhat + beta
    d = d_y * _act_grad(act, pre, _act(act, pre))
    d_ps = np.concatenate([(d * xhat).sum(axis=(0, 1, 3)), d.sum(axis=(0, 1, 3))]).astype(x.dtype)
    if training:
        m1 = d.mean(axis=(0, 1, 3), keepdims=True, dtype=d.dtype)
        m2 = (d * xhat).mean(axis=(0, 1, 3), keepdims=True, dtype=d.dtype)
        d_x = gamma * invstd * (d - m1 - xhat * m2)
    else:
        d_x = gamma * invstd * d
    return d_x, d_ps


def augmenter(x, ps, extra_ch):
    """``AugmenterLayer(Conv((3,3), in => extra; pad=1), 3)``: cat(x, conv(x); dims=3) (src/layers/common.jl:80-92)."""
    return np.concatenate([x, conv2d(x, ps, extra_ch)], axis=2)


def augmenter_vjp(x, ps, d_out):
    cin = x.shape[2]
    d_x, d_ps = conv2d_vjp(x, ps, d_out[:, :, cin:])
    return d_x + d_out[:, :, :cin], d_ps
