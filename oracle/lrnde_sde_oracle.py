"""CPU oracle, SDE side of the hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, in numpy Float32 (Float64 twin by dtype):

* ``NeuralDSDE`` functor                      <- src/layers/neural_sde.jl:1-123
* in-tree SOSRI local-reg step                <- src/perform_step.jl:49-106 (``perform_step_sosri_reg``
                                                 of lrnde_oracle.py is reused verbatim)
* in-tree RKMilCommute / LambaEulerHeun steps <- src/perform_step.jl:108-170, 172-206
* adaptive SOSRI solve loop, RSwM3 rejection sampling, ``sde_determine_initdt``, PI
  controller                                   <- UN-VENDORED StochasticDiffEq 6 /
                                                 DiffEqNoiseProcess 5 ([RECALLED])
* TrackerAdjoint (reverse mode through every accepted step, dt / EEst / noise constant)
                                               <- UN-VENDORED SciMLSensitivity 7

PARITY UNPINNED for everything un-vendored: the reference's Wiener process is seeded from the
global RNG (neural_sde.jl:68-69 passes no seed), so its noise is not reproducible; here both
the oracle and the CUDA path draw from the same counter-based Philox4x32-10 stream
(``philox_normal``), and parity of the loop means "the GPU follows this restatement step for
step".  Every recalled constant lives in ``SDE_CONSTS``.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from .lrnde_oracle import (MLP, SOSRI, LayerOutput, fastpow, perform_step_sosri_reg, rms)

__all__ = ["philox4x32_10", "philox_normal", "SDE_CONSTS", "sde_initdt", "sosri_step",
           "sosri_step_backward", "solve_sosri", "SDESolution", "NeuralDSDE",
           "perform_step_rkmil_reg", "perform_step_lamba_eulerheun_reg", "sosri_reg_backward"]

# [RECALLED] StochasticDiffEq alg_utils.jl defaults for SOSRI (strong order 3/2); SURVEY A.6
SDE_CONSTS = dict(order=1.5, gamma=0.9, qmin=0.2, qmax=1.125, qoldinit=1e-4, delta=1.0 / 6.0,
                  discard_length=1e-15)

# --------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al. 2011), counter-based: the same bits on CPU and GPU
# --------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter words (uint32 arrays); key words are scalars."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + _W0), np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def philox_normal(seed: int, stream: int, draw: int, n: int, dtype=np.float32):
    """n standard normals: element e uses counter (e_lo, e_hi, draw, stream), key = seed;
    Box-Muller in Float64 on the first two output words, rounded to ``dtype``."""
    e = np.arange(n, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10((e & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                                 (e >> np.uint64(32)).astype(np.uint32),
                                 np.full(n, draw, dtype=np.uint32), np.full(n, stream, dtype=np.uint32),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1 = (x0.astype(np.float64) + 0.5) * 2.0 ** -32
    u2 = (x1.astype(np.float64) + 0.5) * 2.0 ** -32
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return z.astype(dtype)


class _Noise:
    """Draw counter shared by the W (stream 0) and Z (stream 1) processes."""

    def __init__(self, seed, shape, dtype, first_draw=0):
        self.seed, self.shape, self.dtype, self.nd = int(seed), shape, dtype, int(first_draw)

    def pair(self):
        n = int(np.prod(self.shape))
        w = philox_normal(self.seed, 0, self.nd, n, self.dtype).reshape(self.shape, order="F")
        z = philox_normal(self.seed, 1, self.nd, n, self.dtype).reshape(self.shape, order="F")
        self.nd += 1
        return w, z


# --------------------------------------------------------------------------
# SOSRI step with everything the reverse pass needs (perform_step.jl:49-106 arithmetic)
# --------------------------------------------------------------------------
def sosri_step(fd, gd, uprev, t, dt, dW, dZ, abstol, reltol, delta):
    u, reg, _, _ = perform_step_sosri_reg(fd, gd, uprev, t, dt, dW, dZ, abstol, reltol, delta)
    T = uprev.dtype.type
    eest = reg / T(dt)
    return u, eest


def _sosri_internals(fd, gd, U, t, dt, dW, dZ):
    T = U.dtype.type
    c = {k: T(v) for k, v in SOSRI.items()}
    t, dt = T(t), T(dt)
    sqdt = np.sqrt(abs(dt))
    chi1 = (dW ** 2 - abs(dt)) / (T(2) * sqdt)
    chi2 = (dW + dZ / np.sqrt(T(3))) / T(2)
    chi3 = (dW ** 3 - T(3) * dW * dt) / (T(6) * dt)
    k1 = fd(U, t); g1 = gd(U, t + c["c11"] * dt)
    H01 = U + dt * c["a021"] * k1 + c["b021"] * chi2 * g1
    H11 = U + dt * c["a121"] * k1 + sqdt * c["b121"] * g1
    k2 = fd(H01, t + c["c02"] * dt); g2 = gd(H11, t + c["c12"] * dt)
    H02 = U + dt * (c["a031"] * k1 + c["a032"] * k2) + chi2 * (c["b031"] * g1 + c["b032"] * g2)
    H12 = U + dt * (c["a131"] * k1 + c["a132"] * k2) + sqdt * (c["b131"] * g1 + c["b132"] * g2)
    k3 = fd(H02, t + c["c03"] * dt); g3 = gd(H12, t + c["c13"] * dt)
    H03 = (U + dt * (c["a041"] * k1 + c["a042"] * k2 + c["a043"] * k3)
           + chi2 * (c["b041"] * g1 + c["b042"] * g2 + c["b043"] * g3))
    H13 = (U + dt * (c["a141"] * k1 + c["a142"] * k2 + c["a143"] * k3)
           + sqdt * (c["b141"] * g1 + c["b142"] * g2 + c["b143"] * g3))
    return dict(c=c, sqdt=sqdt, chi1=chi1, chi2=chi2, chi3=chi3, H=[(U, U), (H01, H11), (H02, H12), (H03, H13)],
                ts=[(t, t + c["c11"] * dt), (t + c["c02"] * dt, t + c["c12"] * dt),
                    (t + c["c03"] * dt, t + c["c13"] * dt), (t + c["c04"] * dt, t + c["c14"] * dt)])


def sosri_step_backward(drift: MLP, diffusion: MLP, ps_f, ps_g, U, t, dt, dW, dZ, ubar,
                        E1bar=None, E2bar=None):
    """Reverse mode through one SOSRI step (dt, noise constant): returns
    (Ubar, d_ps_drift, d_ps_diffusion).  Optional cotangents on E1 / E2 (regulariser)."""
    T = U.dtype.type
    fd = lambda u, tt: drift.f(u, ps_f, tt)
    gd = lambda u, tt: diffusion.f(u, ps_g, tt)
    I = _sosri_internals(fd, gd, U, t, dt, dW, dZ)
    c, sqdt, chi1, chi2, chi3 = I["c"], I["sqdt"], I["chi1"], I["chi2"], I["chi3"]
    dt = T(dt)
    z = np.zeros_like(U)
    E1bar = z if E1bar is None else E1bar
    E2t = ubar + (z if E2bar is None else E2bar)
    kb = [dt * c[f"alpha{i}"] * ubar + dt * E1bar for i in (1, 2, 3, 4)]
    gb = [(chi2 * c[f"beta3{i}"] + chi3 * c[f"beta4{i}"]) * E2t
          + (dW * c[f"beta1{i}"] + chi1 * c[f"beta2{i}"]) * ubar for i in (1, 2, 3, 4)]
    Ub = ubar.copy()
    dpf = np.zeros(drift.nparams, dtype=U.dtype)
    dpg = np.zeros(diffusion.nparams, dtype=U.dtype)
    A0 = {(4, 1): "a041", (4, 2): "a042", (4, 3): "a043", (3, 1): "a031", (3, 2): "a032", (2, 1): "a021"}
    for s in (4, 3, 2):
        H0, H1 = I["H"][s - 1]
        t0s, t1s = I["ts"][s - 1]
        h0b, d0 = drift.vjp(H0, ps_f, t0s, kb[s - 1])
        h1b, d1 = diffusion.vjp(H1, ps_g, t1s, gb[s - 1])
        dpf += d0; dpg += d1
        Ub = Ub + h0b + h1b
        for j in range(1, s):
            a0, a1 = c[A0[(s, j)]], c[A0[(s, j)].replace("a0", "a1")]
            b0, b1 = c[A0[(s, j)].replace("a0", "b0")], c[A0[(s, j)].replace("a0", "b1")]
            kb[j - 1] = kb[j - 1] + dt * (a0 * h0b + a1 * h1b)
            gb[j - 1] = gb[j - 1] + chi2 * b0 * h0b + sqdt * b1 * h1b
    t0s, t1s = I["ts"][0]
    h0b, d0 = drift.vjp(U, ps_f, t0s, kb[0])
    h1b, d1 = diffusion.vjp(U, ps_g, t1s, gb[0])
    dpf += d0; dpg += d1
    Ub = Ub + h0b + h1b
    return Ub, dpf, dpg


def sosri_reg_backward(drift: MLP, diffusion: MLP, ps_f, ps_g, U, t, dt, dW, dZ, abstol, reltol,
                       delta, d_reg):
    """Pullback of reg_val = EEst*dt of perform_step.jl:100-105 w.r.t. the parameters only
    (the integrator state is non-differentiable, neural_sde.jl:41)."""
    T = U.dtype.type
    fd = lambda u, tt: drift.f(u, ps_f, tt)
    gd = lambda u, tt: diffusion.f(u, ps_g, tt)
    I = _sosri_internals(fd, gd, U, t, dt, dW, dZ)
    c, chi1, chi2, chi3 = I["c"], I["chi1"], I["chi2"], I["chi3"]
    ks = [fd(I["H"][i][0], I["ts"][i][0]) for i in range(4)]
    gs = [gd(I["H"][i][1], I["ts"][i][1]) for i in range(4)]
    dtT = T(dt)
    E2 = (chi2 * sum(c[f"beta3{i + 1}"] * gs[i] for i in range(4))
          + chi3 * sum(c[f"beta4{i + 1}"] * gs[i] for i in range(4)))
    u = (U + dtT * sum(c[f"alpha{i + 1}"] * ks[i] for i in range(4)) + E2
         + dW * sum(c[f"beta1{i + 1}"] * gs[i] for i in range(4))
         + chi1 * sum(c[f"beta2{i + 1}"] * gs[i] for i in range(4)))
    E1 = dtT * (ks[0] + ks[1] + ks[2] + ks[3])
    sc = T(abstol) + np.maximum(np.abs(U), np.abs(u)) * T(reltol)
    r = (T(delta) * E1 + E2) / sc
    n = T(u.size)
    eest = np.sqrt(np.sum(r * r, dtype=U.dtype) / n)
    if not eest > 0:
        return np.zeros(drift.nparams, U.dtype), np.zeros(diffusion.nparams, U.dtype)
    rbar = (T(d_reg) * dtT) * r / (n * eest)
    E1bar = T(delta) * rbar / sc
    E2bar = rbar / sc
    scbar = -rbar * r / sc
    ubar = scbar * T(reltol) * np.sign(u) * (np.abs(u) > np.abs(U))
    # E2 enters u as well: sosri_step_backward adds ubar to the E2 cotangent itself
    _, dpf, dpg = sosri_step_backward(drift, diffusion, ps_f, ps_g, U, t, dt, dW, dZ, ubar, E1bar, E2bar)
    return dpf, dpg


# --------------------------------------------------------------------------
# sde_determine_initdt ([RECALLED]; order = 3/2)
# --------------------------------------------------------------------------
def sde_initdt(fd, gd, u0, t, dtmax, abstol, reltol, order=1.5):
    T = u0.dtype.type
    dtmax = T(abs(dtmax))
    sk = T(abstol) + np.abs(u0) * T(reltol)
    d0 = rms(u0 / sk)
    f0 = fd(u0, t)
    g0 = T(3) * gd(u0, t)
    d1 = rms(np.maximum(np.abs(f0 + g0), np.abs(f0 - g0)) / sk)
    dt0 = T(1e-6) if (d0 < T(1e-5) or d1 < T(1e-5)) else T(0.01) * (d0 / d1)
    dt0 = min(dt0, dtmax)
    u1 = u0 + dt0 * f0
    f1 = fd(u1, T(t) + dt0)
    g1 = T(3) * gd(u1, T(t) + dt0)
    dgmax = np.maximum(np.abs(g0 - g1), np.abs(g0 + g1))
    d2 = rms(np.maximum(np.abs(f1 - f0 + dgmax), np.abs(f1 - f0 - dgmax)) / sk) / dt0
    mx = max(d1, d2)
    if mx <= T(1e-15):
        dt1 = max(T(1e-6), dt0 * T(1e-3))
    else:
        dt1 = T(10.0) ** (-(T(2) + np.log10(mx)) / T(order + 0.5))
    return min(T(100) * dt0, dt1, dtmax)


# --------------------------------------------------------------------------
# Adaptive SOSRI loop with RSwM3 (Rackauckas & Nie 2017, Alg. RSwM3)
# --------------------------------------------------------------------------
@dataclass
class SDESolution:
    ts: List = field(default_factory=list)        # accepted step end times (ts[0] = t0)
    us: List = field(default_factory=list)
    steps: List = field(default_factory=list)     # (t, dt, dW, dZ) per accepted step
    log: List = field(default_factory=list)       # (t, dt, EEst, accepted) per attempt
    nf_drift: int = 0
    nf_diffusion: int = 0
    retcode: str = "Success"
    ndraws: int = 0

    def __call__(self, tval):
        """Linear interpolation between accepted steps (SDE dense output is linear)."""
        T = self.us[0].dtype.type
        tval = T(tval)
        for i in range(len(self.ts) - 1):
            if self.ts[i] < tval <= self.ts[i + 1] or (i == 0 and tval == self.ts[0]):
                if tval == self.ts[i + 1]:
                    return self.us[i + 1]
                if tval == self.ts[i]:
                    return self.us[i]
                th = (tval - self.ts[i]) / (self.ts[i + 1] - self.ts[i])
                return (T(1) - th) * self.us[i] + th * self.us[i + 1]
        return self.us[-1]


def _bridge(q, L, K, z, T):
    """N(q*K, (1-q)*q*L) given a standard normal z (variance clamped at 0 against rounding)."""
    return q * K + np.sqrt(max((T(1) - q) * q * L, T(0))) * z


def solve_sosri(fd, gd, u0, t0, t2, *, abstol, reltol, seed, maxiters=1000, pow_mode="fastpow_2023",
                consts=None) -> SDESolution:
    C = dict(SDE_CONSTS, **(consts or {}))
    T = u0.dtype.type
    t0, t2 = T(t0), T(t2)
    beta1, beta2 = T(7.0 / (10.0 * C["order"])), T(2.0 / (5.0 * C["order"]))
    gamma, qmin, qmax = T(C["gamma"]), T(C["qmin"]), T(C["qmax"])
    qold = T(C["qoldinit"])
    delta = C["delta"]
    dtmax = abs(t2 - t0)
    dtmin = max(np.spacing(t0), np.spacing(t2))
    sol = SDESolution(ts=[t0], us=[u0])
    noise = _Noise(seed, u0.shape, u0.dtype)
    t, u = t0, u0
    dt = sde_initdt(fd, gd, u0, t0, dtmax, abstol, reltol, C["order"])
    sol.nf_drift += 2; sol.nf_diffusion += 2
    dt = min(max(dt, dtmin), dtmax, t2 - t)
    S1, S2 = [], []
    z1, z2 = noise.pair()
    dW, dZ = np.sqrt(dt) * z1, np.sqrt(dt) * z2
    it = 0
    while t < t2:
        it += 1
        if it > maxiters:
            sol.retcode = "MaxIters"; break
        if dt <= dtmin and not (t + dt >= t2):
            sol.retcode = "DtLessThanMin"; break
        unew, eest = sosri_step(fd, gd, u, t, dt, dW, dZ, abstol, reltol, delta)
        sol.nf_drift += 4; sol.nf_diffusion += 4
        if not np.isfinite(eest):
            sol.log.append((t, dt, eest, False))
            sol.retcode = "Unstable"; break
        q11 = fastpow(eest, beta1, pow_mode)
        q = q11 / fastpow(qold, beta2, pow_mode)
        q = max(T(1) / qmax, min(T(1) / qmin, q / gamma))
        if eest <= T(1):
            sol.log.append((t, dt, eest, True))
            dtnew = dt / q
            qold = max(eest, T(C["qoldinit"]))
            tnew = t + dt
            if abs(tnew - t2) < T(100) * np.spacing(max(abs(t), abs(t2))):
                tnew = t2
            sol.steps.append((t, dt, dW, dZ))
            sol.ts.append(tnew); sol.us.append(unew)
            t, u = tnew, unew
            if t >= t2:
                break
            dtn = min(max(dtnew, dtmin), dtmax, t2 - t)
            # ---- RSwM3 accept: the increments of the next step come off the future stack
            S2 = []
            dttmp, dW, dZ = T(0), np.zeros_like(u), np.zeros_like(u)
            done = False
            while S1 and not done:
                L, Lw, Lz = S1.pop()
                qtmp = (dtn - dttmp) / L
                if qtmp > T(1):
                    dttmp += L; dW = dW + Lw; dZ = dZ + Lz
                    S2.append((L, Lw, Lz))
                else:
                    z1, z2 = noise.pair()
                    bw, bz = _bridge(qtmp, L, Lw, z1, T), _bridge(qtmp, L, Lz, z2, T)
                    dW = dW + bw; dZ = dZ + bz
                    if (T(1) - qtmp) * L > T(C["discard_length"]):
                        S1.append(((T(1) - qtmp) * L, Lw - bw, Lz - bz))
                    S2.append((qtmp * L, bw, bz))
                    dttmp = dtn
                    done = True
            if not done:
                left = dtn - dttmp
                if left > 0:
                    z1, z2 = noise.pair()
                    bw, bz = np.sqrt(left) * z1, np.sqrt(left) * z2
                    dW = dW + bw; dZ = dZ + bz
                    S2.append((left, bw, bz))
            dt = dtn
        else:
            sol.log.append((t, dt, eest, False))
            dtnew = dt / min(T(1) / qmin, q11 / gamma)
            dtnew = min(max(dtnew, dtmin), dtmax)
            qr = dtnew / dt
            # ---- RSwM3 reject: move whole pieces beyond the kept part back to the future,
            # bridge the piece that straddles the cut
            dttmp, dWtmp, dZtmp = T(0), np.zeros_like(u), np.zeros_like(u)
            while S2:
                L, Lw, Lz = S2.pop()
                if dttmp + L < (T(1) - qr) * dt:
                    dttmp += L; dWtmp = dWtmp + Lw; dZtmp = dZtmp + Lz
                    S1.append((L, Lw, Lz))
                else:
                    S2.append((L, Lw, Lz))
                    break
            dtK = dt - dttmp
            Kw, Kz = dW - dWtmp, dZ - dZtmp
            qK = qr * dt / dtK
            z1, z2 = noise.pair()
            bw, bz = _bridge(qK, dtK, Kw, z1, T), _bridge(qK, dtK, Kz, z2, T)
            cut = (T(1) - qK) * dtK
            if cut > T(C["discard_length"]):
                S1.append((cut, Kw - bw, Kz - bz))
            S2 = [(dtnew, bw, bz)]      # the kept part is re-bridged as one piece
            dW, dZ, dt = bw, bz, dtnew
    sol.ndraws = noise.nd
    return sol


# --------------------------------------------------------------------------
# The other in-tree SDE steps (perform_step.jl:108-170, 172-206), diagonal noise, Ito
# --------------------------------------------------------------------------
def perform_step_rkmil_reg(fd, gd, uprev, t, dt, dW, dZ, abstol, reltol):
    """RKMilCommute (perform_step.jl:108-170).  For diagonal noise J = dW^2/2 (get_iterated_I of
    the commutative algorithm), Ito: J -= |dt|/2 (:120-126).  Quirk kept: the computed drift /
    noise error `tmp` (:162-163) is overwritten (:165); EEst = RMS of (u-uprev)/(abstol +
    max(|uprev|,|u|) reltol) through the 4-argument _calculate_residuals (:218-220)."""
    T = uprev.dtype.type
    t, dt = T(t), T(dt)
    sqdt = np.sqrt(abs(dt))
    J = dW * dW / T(2) - T(0.5) * abs(dt)
    du1 = fd(uprev, t)
    L = gd(uprev, t)
    K = uprev + dt * du1
    tmp = K + sqdt * L
    gtmp = gd(tmp, t)
    Dgj = (gtmp - L) / sqdt
    u = K + L * dW + Dgj * J
    _ = fd(K, t + dt)                       # du2: evaluated, result discarded by the quirk
    resid = (u - uprev) / (T(abstol) + np.maximum(np.abs(uprev), np.abs(u)) * T(reltol))
    EEst = rms(resid)
    return u, EEst * dt, 0, dt


def perform_step_lamba_eulerheun_reg(fd, gd, uprev, t, dt, dW, dZ, abstol, reltol, delta=1.0 / 6.0):
    """LambaEulerHeun (perform_step.jl:172-206), diagonal noise."""
    T = uprev.dtype.type
    t, dt = T(t), T(dt)
    sqdt = np.sqrt(abs(dt))
    du1 = fd(uprev, t)
    K = uprev + dt * du1
    L = gd(uprev, t)
    noise = L * dW
    tmp = K + noise
    gtmp2 = T(0.5) * (L + gd(tmp, t + dt))
    noise2 = gtmp2 * dW
    u = uprev + (dt / T(2)) * (du1 + fd(tmp, t + dt)) + noise2
    du2 = fd(K, t + dt)
    Ed = dt * (du2 - du1) / T(2)
    utilde = uprev + L * sqdt
    ggprime = (gd(utilde, t) - L) / sqdt
    En = ggprime * (dW ** 2) / T(2)
    resid = (T(delta) * Ed + En) / (T(abstol) + np.maximum(np.abs(uprev), np.abs(u)) * T(reltol))
    EEst = np.sqrt(np.sum(resid * resid, dtype=uprev.dtype) / T(u.size))
    return u, EEst * dt, 0, dt


# --------------------------------------------------------------------------
# NeuralDSDE functor (neural_sde.jl:1-123)
# --------------------------------------------------------------------------
class NeuralDSDE:
    """``NeuralDSDE(drift, diffusion; solver=SOSRI(), sensealg=TrackerAdjoint(), tspan=(0f0,1f0),
    regularize=:unbiased, maxiters=1000, kwargs...)`` (neural_sde.jl:11-19).  ``seed`` replaces
    the reference's unseeded Wiener process (see module docstring)."""

    VALID_MODES = ("none", "unbiased", "biased")

    def __init__(self, drift: MLP, diffusion: MLP, *, tspan=(0.0, 1.0), regularize="unbiased",
                 maxiters=1000, dtype=np.float32, pow_mode="fastpow_2023", seed=0, **kwargs):
        if regularize not in self.VALID_MODES:                      # utils.jl:53-58
            raise ValueError(f"regularize must be one of {self.VALID_MODES}")
        self.drift, self.diffusion, self.tspan, self.regularize = drift, diffusion, tspan, regularize
        self.maxiters, self.dtype, self.pow_mode, self.seed = maxiters, np.dtype(dtype), pow_mode, seed
        self.abstol = kwargs.pop("abstol", 1e-2)      # StochasticDiffEq defaults
        self.reltol = kwargs.pop("reltol", 1e-2)
        self.saveat = kwargs.pop("saveat", None)
        self.save_start = kwargs.pop("save_start", None)
        if kwargs:
            raise TypeError(f"unsupported solve kwargs {sorted(kwargs)}")

    def initialstates(self, rng: np.random.Generator):          # neural_sde.jl:22-27
        rng.standard_normal()
        return dict(drift={}, diffusion={}, nfe_drift=-1, nfe_diffusion=-1,
                    reg_val=self.dtype.type(0), rng=copy.deepcopy(rng), training=True)

    def split(self, ps):
        nf = self.drift.nparams
        return ps[:nf], ps[nf:nf + self.diffusion.nparams]     # ps.drift, ps.diffusion (:57,:63)

    def forward(self, x, ps, st):
        T = self.dtype.type
        x = np.asarray(x, dtype=self.dtype)
        ps = np.asarray(ps, dtype=self.dtype)
        pf, pg = self.split(ps)
        fd = lambda u, t: self.drift.f(u, pf, t)
        gd = lambda u, t: self.diffusion.f(u, pg, t)
        t0, t2 = T(self.tspan[0]), T(self.tspan[1])
        mode = self.regularize if st["training"] else "none"
        sol = solve_sosri(fd, gd, x, t0, t2, abstol=self.abstol, reltol=self.reltol, seed=self.seed,
                          maxiters=self.maxiters, pow_mode=self.pow_mode)
        nfd, nfg = sol.nf_drift, sol.nf_diffusion

        def saves(times):
            times = [T(s) for s in times]
            if self.save_start and t0 not in times:
                times = [t0] + times
            tl = sol.ts[-1]
            return LayerOutput(times, [sol(min(s, tl)) for s in times])

        def all_steps():
            start = 0 if (self.save_start is None or self.save_start) else 1
            return LayerOutput(list(sol.ts[start:]), list(sol.us[start:]))

        if mode == "none":
            out = saves([t2] if self.saveat is None else self.saveat)
            st2 = dict(st, nfe_drift=nfd, nfe_diffusion=nfg, reg_val=T(0))
            return out, st2, dict(sol=sol, mode=mode, out_full=out, x=x, ps=ps)
        rng = copy.deepcopy(st["rng"])
        if mode == "unbiased":
            t1 = T(rng.random(dtype=np.float32)) * (t2 - t0) + t0            # :89
            if self.saveat is None:
                out_full = out = saves([t1, t2])
            else:
                out_full = saves(list(self.saveat) + [t1])
                keep = [i for i, s in enumerate(out_full.t) if s != t1]
                out = LayerOutput([out_full.t[i] for i in keep], [out_full.u[i] for i in keep])
        else:
            out_full = out = all_steps() if self.saveat is None else saves(self.saveat)
            cand = out_full.t[:-1]                                        # :112
            # rand(rng, sol.t[1:end-1]) restated as index = floor(u01 * ncand) (the convention of
            # lrnde_sde_opts.u01, include/lrnde.h)
            u01 = float(rng.random(dtype=np.float32))
            t1 = cand[min(len(cand) - 1, int(np.floor(np.float32(u01) * np.float32(len(cand)))))] if cand else t0
        u1 = sol(t1)
        # _get_dsde_integrator -> init(SDEProblem(dudt, g, u1, (t1,t2), ps), SOSRI()): auto dt,
        # fresh Wiener increments for that dt (non-differentiable, :41)
        dt = sde_initdt(fd, gd, u1, t1, abs(t2 - t1), self.abstol, self.reltol)
        dtmin = max(np.spacing(t1), np.spacing(t2))
        dt = min(max(dt, dtmin), abs(t2 - t1)) if t2 > t1 else dtmin
        nz = _Noise(self.seed, u1.shape, u1.dtype, first_draw=sol.ndraws)
        z1, z2 = nz.pair()
        dW, dZ = np.sqrt(dt) * z1, np.sqrt(dt) * z2
        _, reg, _, _ = perform_step_sosri_reg(fd, gd, u1, t1, dt, dW, dZ, self.abstol, self.reltol,
                                              SDE_CONSTS["delta"])
        nfd += 2 + 4; nfg += 2 + 4       # the closures count the integrator's evaluations too (:44-65)
        st2 = dict(st, nfe_drift=nfd, nfe_diffusion=nfg, reg_val=T(reg), rng=rng)
        aux = dict(sol=sol, mode=mode, t1=t1, u1=u1, dt_reg=dt, dW_reg=dW, dZ_reg=dZ, out_full=out_full,
                   x=x, ps=ps)
        return out, st2, aux

    def __call__(self, x, ps, st):
        out, st2, _ = self.forward(x, ps, st)
        return out, st2

    def backward(self, aux, d_us, d_reg, ps):
        """TrackerAdjoint: reverse mode through every accepted step and through the linear
        interpolation that produced each saved state; pullback of reg_val w.r.t. ps only."""
        ps = np.asarray(ps, dtype=self.dtype)
        T = self.dtype.type
        pf, pg = self.split(ps)
        sol = aux["sol"]
        out_full = aux["out_full"]
        if aux["mode"] == "unbiased" and self.saveat is not None:
            it = iter(d_us)
            d_full = [None if s == aux["t1"] else next(it) for s in out_full.t]
        else:
            d_full = list(d_us)
        n = len(sol.ts)
        cot = [np.zeros_like(sol.us[0]) for _ in range(n)]         # cotangent on each step state
        tl = sol.ts[-1]
        for s, d in zip(out_full.t, d_full):
            if d is None:
                continue
            s = min(T(s), tl)
            d = np.asarray(d, dtype=self.dtype)
            for i in range(n - 1):
                if sol.ts[i] < s <= sol.ts[i + 1] or (i == 0 and s == sol.ts[0]):
                    if s == sol.ts[i + 1]:
                        cot[i + 1] += d
                    elif s == sol.ts[i]:
                        cot[i] += d
                    else:
                        th = (s - sol.ts[i]) / (sol.ts[i + 1] - sol.ts[i])
                        cot[i] += (T(1) - th) * d
                        cot[i + 1] += th * d
                    break
            else:
                cot[-1] += d
        dpf = np.zeros(self.drift.nparams, self.dtype)
        dpg = np.zeros(self.diffusion.nparams, self.dtype)
        ub = cot[n - 1]
        for i in range(n - 2, -1, -1):
            t, dt, dW, dZ = sol.steps[i]
            ub, a, b = sosri_step_backward(self.drift, self.diffusion, pf, pg, sol.us[i], t, dt, dW, dZ, ub)
            dpf += a; dpg += b
            ub = ub + cot[i]
        if aux["mode"] != "none" and d_reg is not None and d_reg != 0:
            a, b = sosri_reg_backward(self.drift, self.diffusion, pf, pg, aux["u1"], aux["t1"], aux["dt_reg"],
                                      aux["dW_reg"], aux["dZ_reg"], self.abstol, self.reltol,
                                      SDE_CONSTS["delta"], d_reg)
            dpf += a; dpg += b
        return ub, np.concatenate([dpf, dpg])
