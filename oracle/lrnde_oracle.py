"""CPU restatement (numpy, Float32 by default) of the LocalRegNeuralDE hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED.

What is restated, and from where (all paths relative to /root/reference):

* ``perform_step_tsit5_reg``        <- src/perform_step.jl:3-32   (arithmetic order kept)
* ``reg_error_estimate``            <- src/perform_step.jl:34-38
* ``reg_stiffness_estimate``        <- src/perform_step.jl:40-47
* ``calculate_residuals`` / ``rms`` <- src/perform_step.jl:208-220
* ``perform_step_sosri_reg``        <- src/perform_step.jl:49-106
* ``NeuralODE`` (modes, t1, saveat, nfe, state tuple)
                                    <- src/layers/neural_ode.jl:1-118
* ``MLP`` (TDChain time row appended before *every* layer; ArrayAndTime)
                                    <- src/layers/common.jl:10-45, src/utils.jl:12-23
* ``diffeqsol_to_array/_timeseries``<- src/utils.jl:25-46

The arithmetic that lives in UN-VENDORED Julia dependencies (no Manifest.toml in
the reference, compat bounds only: OrdinaryDiffEq 6, DiffEqBase 6,
SciMLSensitivity 7, Lux 0.4.41-0.5, NNlib 0.8-0.9; Project.toml:31-53) is restated
from the published algorithms as recorded in SURVEY.md Appendix A:

* ``TSIT5`` tableau + free interpolant      (OrdinaryDiffEq Tsit5ConstantCache)
* ``ode_initdt``                            (OrdinaryDiffEq ode_determine_initdt, oop)
* ``solve_tsit5`` loopheader!/perform_step!/loopfooter!, PI controller,
  ``fastpow`` (DiffEqBase 2023 Float32 approximation), tstop clamping/snapping
* ``adjoint_backward``                      (SciMLSensitivity InterpolatingAdjoint
  + ZygoteVJP: augmented [lambda; mu] integrated t2->t0 with the same Tsit5,
  lambda jumps at the saved times which are also tstops, error norm over all of z)
* Lux ``Dense`` / activations               (Lux/NNlib)

Because neither Julia nor any golden vector exists here, this file IS the
definition the CUDA path is held to; it is self-validated in tests/ (tableau
order conditions, convergence order, analytic ODEs, adjoint vs finite
differences in Float64, the reference's own 9 property tests).

Layout: logical arrays are ``(features, batch)`` like the reference; flat
buffers exchanged with libLRNDE.so are the column-major raveling of those
(``a.ravel(order="F")``), parameters in ComponentArray order
``layer_1.weight[out x in(+1)]``, ``layer_1.bias[out]``, ``layer_2.weight`` ...
"""
from __future__ import annotations

import copy
import math
import struct
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

__all__ = [
    "TSIT5", "Tableau", "Dense", "MLP", "fastpow", "rms", "calculate_residuals",
    "ode_initdt", "solve_tsit5", "ODESolution", "perform_step_tsit5_reg",
    "reg_error_estimate", "reg_stiffness_estimate", "reg_step_backward",
    "adjoint_backward", "NeuralODE", "LayerOutput", "diffeqsol_to_array",
    "diffeqsol_to_timeseries", "glorot_uniform_params", "SOSRI", "perform_step_sosri_reg",
    "RETCODE_SUCCESS", "RETCODE_MAXITERS", "RETCODE_DTMIN", "RETCODE_UNSTABLE",
    "mnist_ode_model", "interp_step",
]

RETCODE_SUCCESS, RETCODE_MAXITERS, RETCODE_DTMIN, RETCODE_UNSTABLE = 0, 1, 2, 3

# --------------------------------------------------------------------------
# Tsit5 tableau + interpolant
# --------------------------------------------------------------------------
_TSIT5_F64 = dict(
    c1=0.161, c2=0.327, c3=0.9, c4=0.9800255409045097,
    a21=0.161,
    a31=-0.008480655492356989, a32=0.335480655492357,
    a41=2.8971530571054935, a42=-6.359448489975075, a43=4.3622954328695815,
    a51=5.325864828439257, a52=-11.748883564062828, a53=7.4955393428898365,
    a54=-0.09249506636175525,
    a61=5.86145544294642, a62=-12.92096931784711, a63=8.159367898576159,
    a64=-0.071584973281401, a65=-0.028269050394068383,
    a71=0.09646076681806523, a72=0.01, a73=0.4798896504144996, a74=1.379008574103742,
    a75=-3.290069515436081, a76=2.324710524099774,
    btilde1=-0.00178001105222577714, btilde2=-0.0008164344596567469,
    btilde3=0.007880878010261995, btilde4=-0.1447110071732629,
    btilde5=0.5823571654525552, btilde6=-0.45808210592918697,
    btilde7=0.015151515151515152,
    r11=1.0, r12=-2.763706197274826, r13=2.9132554618219126, r14=-1.0530884977290216,
    r22=0.13169999999999998, r23=-0.2234, r24=0.1017,
    r32=3.9302962368947516, r33=-5.941033872131505, r34=2.490627285651253,
    r42=-12.411077166933676, r43=30.33818863028232, r44=-16.548102889244902,
    r52=37.50931341651104, r53=-88.1789048947664, r54=47.37952196281928,
    r62=-27.896526289197286, r63=65.09189467479366, r64=-34.87065786149661,
    r72=1.5, r73=-4.0, r74=2.5,
)


class Tableau:
    """Tsit5 constants rounded to ``dtype`` (perform_step.jl:6-8 extracts them at
    T=T2=Float32)."""

    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        for k, v in _TSIT5_F64.items():
            setattr(self, k, self.dtype.type(v))
        T = self.dtype.type
        self.c = [T(0), self.c1, self.c2, self.c3, self.c4, T(1), T(1)]
        self.a = [
            [],
            [self.a21],
            [self.a31, self.a32],
            [self.a41, self.a42, self.a43],
            [self.a51, self.a52, self.a53, self.a54],
            [self.a61, self.a62, self.a63, self.a64, self.a65],
            [self.a71, self.a72, self.a73, self.a74, self.a75, self.a76],
        ]
        self.btilde = [self.btilde1, self.btilde2, self.btilde3, self.btilde4,
                       self.btilde5, self.btilde6, self.btilde7]

    def interp_weights(self, theta):
        """b_i(theta), i=1..7, of the free 4th-order interpolant."""
        T = self.dtype.type
        th = T(theta)
        th2 = th * th
        b1 = th * (self.r11 + th * (self.r12 + th * (self.r13 + th * self.r14)))
        out = [b1]
        for i in range(2, 8):
            r2, r3, r4 = (getattr(self, f"r{i}{j}") for j in (2, 3, 4))
            out.append(th2 * (r2 + th * (r3 + th * r4)))
        return out


TSIT5 = Tableau(np.float32)
_TABLEAUS = {np.dtype(np.float32): TSIT5, np.dtype(np.float64): Tableau(np.float64)}


def _tab(dtype) -> Tableau:
    return _TABLEAUS[np.dtype(dtype)]


# --------------------------------------------------------------------------
# Activations and the (time-dependent) MLP dynamics
# --------------------------------------------------------------------------
_SQRT_2_OVER_PI = 0.7978845608028654


def _act(name: str, x):
    T = x.dtype.type
    if name == "identity":
        return x
    if name == "tanh":
        return np.tanh(x)
    if name == "sigmoid":
        return T(1) / (T(1) + np.exp(-x))
    if name == "relu":
        return np.maximum(x, T(0))
    if name == "gelu":  # NNlib tanh form
        inner = T(_SQRT_2_OVER_PI) * (x + T(0.044715) * x * x * x)
        return T(0.5) * x * (T(1) + np.tanh(inner))
    raise ValueError(f"unknown activation {name}")


def _act_grad(name: str, pre, out):
    """d act / d pre, given the pre-activation and the activation output."""
    T = pre.dtype.type
    if name == "identity":
        return np.ones_like(pre)
    if name == "tanh":
        return T(1) - out * out
    if name == "sigmoid":
        return out * (T(1) - out)
    if name == "relu":
        return (pre > 0).astype(pre.dtype)
    if name == "gelu":
        x = pre
        inner = T(_SQRT_2_OVER_PI) * (x + T(0.044715) * x * x * x)
        th = np.tanh(inner)
        dinner = T(_SQRT_2_OVER_PI) * (T(1) + T(3 * 0.044715) * x * x)
        return T(0.5) * (T(1) + th) + T(0.5) * x * (T(1) - th * th) * dinner
    raise ValueError(name)


@dataclass(frozen=True)
class Dense:
    """Lux ``Dense(in => out, act)``: ``act.(W*x .+ b)``.  ``in_dims``
    excludes the time row a TDChain appends."""
    in_dims: int
    out_dims: int
    act: str = "identity"


@dataclass
class MLP:
    """``Chain(Dense...)`` optionally wrapped in ``TDChain`` (common.jl:2-45): when
    ``time_dependent`` the scalar ``t`` is appended as an extra last input row before
    every layer (common.jl:19-33), i.e. every weight has ``in_dims+1`` columns.
    ``input_act`` models a leading ``Base.Fix1(broadcast, tanh)`` (construct.jl:235)."""
    layers: List[Dense]
    time_dependent: bool = True
    input_act: Optional[str] = None

    def __post_init__(self):
        td = 1 if self.time_dependent else 0
        off = 0
        self.offsets = []
        for L in self.layers:
            w_off = off
            off += L.out_dims * (L.in_dims + td)
            b_off = off
            off += L.out_dims
            self.offsets.append((w_off, b_off))
        self.nparams = off
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            assert a.out_dims == b.in_dims
        assert self.layers[0].in_dims == self.layers[-1].out_dims, "f: R^D -> R^D"

    @property
    def state_dims(self) -> int:
        return self.layers[0].in_dims

    def unpack(self, ps):
        td = 1 if self.time_dependent else 0
        out = []
        for L, (wo, bo) in zip(self.layers, self.offsets):
            W = ps[wo:wo + L.out_dims * (L.in_dims + td)].reshape(
                (L.out_dims, L.in_dims + td), order="F")
            b = ps[bo:bo + L.out_dims]
            out.append((W, b))
        return out

    # f(u, p, t): the ``dudt`` closure of neural_ode.jl:45-48
    def f(self, u, ps, t, cache: Optional[list] = None):
        T = u.dtype.type
        x = u
        if self.input_act is not None:
            x0 = x
            x = _act(self.input_act, x)
            if cache is not None:
                cache.append(("in", x0, x))
        for L, (W, b) in zip(self.layers, self.unpack(ps)):
            if self.time_dependent:
                xin = np.concatenate([x, np.full((1, x.shape[1]), T(t), dtype=x.dtype)], axis=0)
            else:
                xin = x
            pre = W @ xin + b[:, None]
            y = _act(L.act, pre)
            if cache is not None:
                cache.append((xin, pre, y))
            x = y
        return x

    def vjp(self, u, ps, t, lam):
        """(J_u^T lam, J_p^T lam) of f at (u, t) -- what Zygote.pullback((u,p)->f(u,p,t))
        returns in SciMLSensitivity's ZygoteVJP."""
        cache: list = []
        self.f(u, ps, t, cache)
        dps = np.zeros(self.nparams, dtype=u.dtype)
        g = lam
        layer_caches = cache[1:] if self.input_act is not None else cache
        for L, (wo, bo), (W, _b), (xin, pre, y) in reversed(list(zip(
                self.layers, self.offsets, self.unpack(ps), layer_caches))):
            d = g * _act_grad(L.act, pre, y)
            dW = d @ xin.T
            dps[wo:wo + dW.size] = dW.ravel(order="F")
            dps[bo:bo + L.out_dims] = d.sum(axis=1)
            g = W.T @ d
            if self.time_dependent:
                g = g[:-1, :]
        if self.input_act is not None:
            _, x0, x = cache[0]
            g = g * _act_grad(self.input_act, x0, x)
        return g, dps


def glorot_uniform_params(model: MLP, rng: np.random.Generator, dtype=np.float32):
    """Lux default init: Glorot-uniform weights, zero bias."""
    td = 1 if model.time_dependent else 0
    ps = np.zeros(model.nparams, dtype=dtype)
    for L, (wo, bo) in zip(model.layers, model.offsets):
        fan_in, fan_out = L.in_dims + td, L.out_dims
        a = math.sqrt(6.0 / (fan_in + fan_out))
        W = rng.uniform(-a, a, size=(fan_out, fan_in)).astype(dtype)
        ps[wo:wo + W.size] = W.ravel(order="F")
    return ps


def mnist_ode_model(D=784, H=100) -> MLP:
    """experiments/src/construct.jl:180-189 with config.jl:26-28 defaults."""
    return MLP([Dense(D, H, "tanh"), Dense(H, D, "identity")], time_dependent=True)


# --------------------------------------------------------------------------
# Norms / residuals (perform_step.jl:208-220)
# --------------------------------------------------------------------------
def rms(x):
    """``_internalnorm``: sqrt(mean(abs2, x)) (perform_step.jl:208), same as the
    OrdinaryDiffEq default norm sqrt(sum(abs2,u)/length(u))."""
    T = x.dtype.type
    return np.sqrt(np.sum(x * x, dtype=x.dtype) / T(x.size))


def calculate_residuals(utilde, u0, u1, alpha, rho):
    """perform_step.jl:210-212."""
    return utilde / (alpha + np.maximum(np.abs(u0), np.abs(u1)) * rho)


# --------------------------------------------------------------------------
# DiffEqBase.fastpow, 2023 Float32 version
# --------------------------------------------------------------------------
def _f2u(x) -> int:
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def _u2f(b: int) -> np.float32:
    return np.float32(struct.unpack("<f", struct.pack("<I", b & 0xFFFFFFFF))[0])


def _fastlog2(x) -> np.float32:
    f = np.float32
    bits = _f2u(f(x))
    e = (bits & 0x7F800000) >> 23
    if bits & 0x00400000:
        s = _u2f((bits & 0x007FFFFF) | 0x3F000000)
        fe = f(e) - f(126.0)
    else:
        s = _u2f((bits & 0x007FFFFF) | 0x3F800000)
        fe = f(e) - f(127.0)
    s = s - f(1.0)
    return fe + s * (f(0.338953) * s + f(2.198599)) / (s + f(1.523692))


def _fastpow2(x) -> np.float32:
    f = np.float32
    x = f(x)
    offset = f(1.0) if x < 0 else f(0.0)
    clipp = f(-126.0) if x < f(-126.0) else x
    w = f(int(clipp))  # trunc toward zero
    z = clipp - w + offset
    v = f(1 << 23) * (clipp + f(121.2740575) + f(27.7280233) / (f(4.84252568) - z)
                      - f(1.49012907) * z)
    return _u2f(int(v))


def fastpow(x, y, mode: str = "fastpow_2023"):
    """x^y as the controller computes it.  ``fastpow_2023`` is the Float32
    approximation DiffEqBase used in the era the reference targets; ``exact`` is
    libm pow (later versions / Float64)."""
    if mode == "exact" or np.dtype(type(x)) == np.float64:
        return type(x)(float(x) ** float(y)) if isinstance(x, np.floating) else float(x) ** float(y)
    x = np.float32(x)
    if x == 0:
        return np.float32(0)
    if np.isinf(x) and np.isinf(np.float32(y)):
        return np.float32(np.inf)
    return _fastpow2(np.float32(y) * _fastlog2(x))


# --------------------------------------------------------------------------
# Initial step size (out-of-place ode_determine_initdt)
# --------------------------------------------------------------------------
def ode_initdt(f, u0, t, tdir, dtmax, abstol, reltol, f0=None, order=5, dtmin=None):
    """Returns (dt, number of extra f evaluations).  ``f(u, t)``.  The library
    re-evaluates f0 here even though fsalfirst already holds it (nf += 2)."""
    T = u0.dtype.type
    if dtmin is None:
        dtmin = T(np.finfo(u0.dtype).eps)
    dtmax = T(abs(dtmax))
    sk = T(abstol) + np.abs(u0) * T(reltol)
    d0 = rms(u0 / sk)
    if f0 is None:
        f0 = f(u0, t)
    if not np.all(np.isfinite(f0)):
        # library warns and returns tdir*dtmin; keep it total.
        return T(tdir) * dtmin, 2
    d1 = rms(f0 / sk)
    if d0 < T(1e-5) or d1 < T(1e-5):
        dt0 = T(1e-6)
    else:
        dt0 = T(0.01) * (d0 / d1)
    dt0 = min(dt0, dtmax)
    dt0_s = T(tdir) * dt0
    u1 = u0 + dt0_s * f0
    f1 = f(u1, T(t) + dt0_s)
    if np.array_equal(f0, f1):
        return T(tdir) * max(dtmin, T(100) * dt0), 2
    d2 = rms((f1 - f0) / sk) / dt0
    mx = max(d1, d2)
    if mx <= T(1e-15):
        dt1 = max(T(1e-6), dt0 * T(1e-3))
    else:
        dt1 = T(10.0) ** (-(T(2) + np.log10(mx)) / T(order))
    return T(tdir) * max(dtmin, min(T(100) * dt0, dt1, dtmax)), 2


# --------------------------------------------------------------------------
# Adaptive Tsit5 solve
# --------------------------------------------------------------------------
@dataclass
class ODESolution:
    """Dense solution: accepted step times ``ts`` (ts[0]=t0), states ``us`` and for
    step n (ts[n] -> ts[n+1]) the seven stage derivatives ``ks[n]`` and ``dts[n]``."""
    ts: list
    us: list
    ks: list
    dts: list
    step_log: list            # (t, dt, EEst, accepted) for EVERY attempted step
    nf: int
    naccept: int
    nreject: int
    retcode: int
    tdir: int
    dtype: np.dtype
    t: list = field(default_factory=list)   # saved times (saveat)
    u: list = field(default_factory=list)   # saved states

    def locate(self, tval):
        """Interval index n with ts[n] < tval <= ts[n+1] (forward; mirrored for
        tdir<0), n=0 when tval == ts[0] -- OrdinaryDiffEq ode_interpolation's
        ``continuity=:left`` search."""
        n_steps = len(self.dts)
        d = self.tdir
        lo, hi = 1, n_steps          # searchsortedfirst over ts[1:]
        while lo < hi:
            mid = (lo + hi) // 2
            if d * self.ts[mid] >= d * tval:
                hi = mid
            else:
                lo = mid + 1
        return lo - 1

    def __call__(self, tval, exact_hit_stored=False):
        T = self.dtype.type
        tval = T(tval)
        if exact_hit_stored:
            for n, tn in enumerate(self.ts):
                if tn == tval:
                    return self.us[n]
        n = self.locate(tval)
        dt = self.ts[n + 1] - self.ts[n]
        theta = (tval - self.ts[n]) / dt
        return interp_step(self.us[n], self.ks[n], dt, theta)


def interp_step(u0, ks, dt, theta):
    """Tsit5 free interpolant u(theta) = u0 + dt*sum_i b_i(theta) k_i."""
    tab = _tab(u0.dtype)
    b = tab.interp_weights(theta)
    acc = ks[0] * b[0]
    for i in range(1, 7):
        acc = acc + ks[i] * b[i]
    return u0 + u0.dtype.type(dt) * acc


def tsit5_stages(f, uprev, k1, t, dt):
    """k2..k7, u, utilde of one Tsit5 step; arithmetic order of perform_step.jl:10-27."""
    tab = _tab(uprev.dtype)
    T = uprev.dtype.type
    t = T(t)
    dt = T(dt)
    a = dt * tab.a21
    k2 = f(uprev + a * k1, t + tab.c1 * dt)
    k3 = f(uprev + dt * (tab.a31 * k1 + tab.a32 * k2), t + tab.c2 * dt)
    k4 = f(uprev + dt * (tab.a41 * k1 + tab.a42 * k2 + tab.a43 * k3), t + tab.c3 * dt)
    k5 = f(uprev + dt * (tab.a51 * k1 + tab.a52 * k2 + tab.a53 * k3 + tab.a54 * k4),
           t + tab.c4 * dt)
    g6 = uprev + dt * (tab.a61 * k1 + tab.a62 * k2 + tab.a63 * k3 + tab.a64 * k4
                       + tab.a65 * k5)
    k6 = f(g6, t + dt)
    u = uprev + dt * (tab.a71 * k1 + tab.a72 * k2 + tab.a73 * k3 + tab.a74 * k4
                      + tab.a75 * k5 + tab.a76 * k6)
    k7 = f(u, t + dt)
    utilde = dt * (tab.btilde1 * k1 + tab.btilde2 * k2 + tab.btilde3 * k3
                   + tab.btilde4 * k4 + tab.btilde5 * k5 + tab.btilde6 * k6
                   + tab.btilde7 * k7)
    return [k1, k2, k3, k4, k5, k6, k7], u, utilde, g6


def solve_tsit5(f: Callable, u0, t0, tend, *, abstol, reltol, maxiters=1000,
                tstops: Sequence = (), on_tstop: Optional[Callable] = None,
                pow_mode: str = "fastpow_2023", init_jump: Optional[Callable] = None,
                dt0=None) -> ODESolution:
    """OrdinaryDiffEq ``solve(prob, Tsit5(); abstol, reltol, maxiters, tstops)`` on
    a constant cache, dense output kept.  ``f(u, t)``.  Works for either time
    direction.  ``tstops`` strictly inside the span are honoured (dt clamped to land
    on them); ``on_tstop(t, u)`` may return a modified u (a discrete callback), after
    which fsalfirst is re-evaluated (nf += 1).  Solver failures are retcodes, never
    exceptions (reference: utils.jl:61 marks check_error! non-differentiable and
    carries on)."""
    dtype = u0.dtype
    T = dtype.type
    t0 = T(t0)
    tend = T(tend)
    tdir = 1 if tend >= t0 else -1
    eps = T(np.finfo(dtype).eps)
    # controller defaults
    qmin, qmax, gamma = T(1) / T(5), T(10), T(9) / T(10)
    beta2, beta1 = T(2) / T(5 * 5), T(7) / T(10 * 5)
    qoldinit = T(1e-4)
    dtmax = abs(tend - t0)
    dtmin = T(max(np.spacing(abs(t0)), np.spacing(abs(tend))))   # prob2dtmin

    stops = sorted({T(s) for s in tstops if tdir * t0 < tdir * T(s) < tdir * tend},
                   key=lambda s: tdir * s)
    stops.append(tend)

    u = u0
    if init_jump is not None:
        unew = init_jump(t0, u)
        if unew is not None:
            u = unew
    nf = 0
    k1 = f(u, t0)
    nf += 1
    if dt0 is None:
        dt, extra = ode_initdt(f, u, t0, tdir, dtmax, abstol, reltol, order=5, dtmin=dtmin)
        nf += extra
    else:
        dt = T(dt0)

    sol = ODESolution(ts=[t0], us=[u], ks=[], dts=[], step_log=[], nf=0, naccept=0,
                      nreject=0, retcode=RETCODE_SUCCESS, tdir=tdir, dtype=dtype)
    t = t0
    qold = qoldinit
    q11 = T(1)
    it = 0
    accepted_prev = True
    dtpropose = dt
    first = True
    while stops:
        tstop = stops[0]
        while tdir * t < tdir * tstop:
            # ---- loopheader!
            if not first:
                if accepted_prev:
                    dt = dtpropose
                else:
                    dt = dt / min(T(1) / qmin, q11 / gamma)
            first = False
            it += 1
            dt = T(tdir) * min(abs(dt), dtmax)
            dt = T(tdir) * max(abs(dt), dtmin)
            clamped = abs(tstop - t) <= abs(dt)
            dt = T(tdir) * min(abs(dt), abs(tstop - t))
            # ---- check_error!
            if it > maxiters:
                sol.retcode = RETCODE_MAXITERS
                break
            if abs(dt) <= dtmin and not clamped:
                sol.retcode = RETCODE_DTMIN
                break
            if np.any(np.isnan(u)):
                sol.retcode = RETCODE_UNSTABLE
                break
            # ---- perform_step!
            ks, unew, utilde, _ = tsit5_stages(f, u, k1, t, dt)
            nf += 6
            EEst = rms(calculate_residuals(utilde, u, unew, T(abstol), T(reltol)))
            # ---- loopfooter!: PI controller
            if EEst == 0:
                q = T(1) / qmax
            else:
                q11 = fastpow(EEst, beta1, pow_mode)
                q = q11 / fastpow(qold, beta2, pow_mode)
                q = max(T(1) / qmax, min(T(1) / qmin, q / gamma))
            accept = bool(EEst <= 1)    # NaN -> reject
            sol.step_log.append((t, dt, EEst, accept))
            if accept:
                sol.naccept += 1
                if T(1) <= q <= T(1):       # qsteady_min = qsteady_max = 1
                    q = T(1)
                qold = max(EEst, qoldinit)
                dtnew = dt / q
                ttmp = t + dt
                if abs(ttmp - tstop) < T(100) * np.spacing(max(abs(t), abs(tstop))).astype(dtype):
                    ttmp = tstop
                dtpropose = T(tdir) * max(min(abs(dtnew), dtmax), dtmin)
                sol.ks.append(ks)
                sol.dts.append(ttmp - t)
                t = ttmp
                u = unew
                k1 = ks[6]
                sol.ts.append(t)
                sol.us.append(u)
            else:
                sol.nreject += 1
            accepted_prev = accept
        if sol.retcode != RETCODE_SUCCESS:
            break
        stops.pop(0)
        if stops and on_tstop is not None:
            unew = on_tstop(t, u)
            if unew is not None:
                u = unew
                sol.us[-1] = u   # the stored left state of the next interval is post-jump
                k1 = f(u, t)
                nf += 1
    sol.nf = nf
    return sol


# --------------------------------------------------------------------------
# The local-regularisation step (src/perform_step.jl:3-47)
# --------------------------------------------------------------------------
def reg_error_estimate(utilde, uprev, u, abstol, reltol, dt):
    """perform_step.jl:34-38."""
    T = u.dtype.type
    r = calculate_residuals(utilde, uprev, u, T(abstol), T(reltol))
    return np.sqrt(np.sum(r * r, dtype=u.dtype) / T(u.size)) * T(dt)


def reg_stiffness_estimate(k7, k6, g7, g6):
    """perform_step.jl:40-47 (3.5068f0 = alg_stability_size(Tsit5()))."""
    T = k7.dtype.type
    d = g7 - g6
    den = np.sqrt(np.mean(d * d, dtype=k7.dtype))
    if den == 0:
        return T(0)
    dk = k7 - k6
    return abs(np.sqrt(np.mean(dk * dk, dtype=k7.dtype)) / (den + T(np.finfo(k7.dtype).eps))) \
        / T(3.5068)


def perform_step_tsit5_reg(f, uprev, k1, t, dt, abstol, reltol, reg_type="error_estimate",
                           nf_init=3):
    """``_perform_step(integrator, ::Tsit5ConstantCache, p, Val(reg_type))``
    (perform_step.jl:3-32).  Returns (u, reg_val, nfe, dt, stages) with
    nfe = 6 + integrator.sol.destats.nf (:31)."""
    ks, u, utilde, g6 = tsit5_stages(f, uprev, k1, t, dt)
    if reg_type == "error_estimate":
        reg = reg_error_estimate(utilde, uprev, u, abstol, reltol, dt)
    elif reg_type == "stiffness_estimate":
        reg = reg_stiffness_estimate(ks[6], ks[5], u, g6)
    else:
        raise ValueError("regularize_type must be one of (:error_estimate, :stiffness_estimate)")
    return u, reg, 6 + nf_init, dt, dict(ks=ks, utilde=utilde, g6=g6, u=u)


def reg_step_backward(model: MLP, ps, uprev, k1, t, dt, abstol, reltol, reg_type, d_reg):
    """Reverse pass of ``_perform_step`` w.r.t. ``ps`` ONLY: uprev=u(t1), k1 and dt
    come from the non-differentiable integrator (neural_ode.jl:40, utils.jl:60;
    pinned by test/runtests.jl:127-131)."""
    dtype = uprev.dtype
    T = dtype.type
    tab = _tab(dtype)
    t = T(t)
    dt = T(dt)
    f = lambda u, tt: model.f(u, ps, tt)
    ks, u, utilde, g6 = tsit5_stages(f, uprev, k1, t, dt)
    n = T(u.size)
    dk = [np.zeros_like(u) for _ in range(7)]
    du = np.zeros_like(u)
    dg6 = np.zeros_like(u)
    if reg_type == "error_estimate":
        denom = T(abstol) + np.maximum(np.abs(uprev), np.abs(u)) * T(reltol)
        r = utilde / denom
        ss = np.sum(r * r, dtype=dtype)
        root = np.sqrt(ss / n)
        if root == 0:
            return np.zeros(model.nparams, dtype=dtype)
        # reg = sqrt(ss/n)*dt
        dr = T(d_reg) * dt * r / (n * root)
        dutilde = dr / denom
        du = du + (-dr * utilde / (denom * denom)) * T(reltol) * np.sign(u) * (np.abs(u) > np.abs(uprev))
        for i in range(1, 7):
            dk[i] = dk[i] + dt * tab.btilde[i] * dutilde
    else:
        d = u - g6
        den = np.sqrt(np.mean(d * d, dtype=dtype))
        if den == 0:
            return np.zeros(model.nparams, dtype=dtype)
        kk = ks[6] - ks[5]
        num = np.sqrt(np.mean(kk * kk, dtype=dtype))
        eps = T(np.finfo(dtype).eps)
        # est = num/(den+eps)/3.5068
        c = T(d_reg) / T(3.5068)
        dnum = c / (den + eps)
        dden = -c * num / ((den + eps) * (den + eps))
        if num != 0:
            dkk = dnum * kk / (n * num)
            dk[6] = dk[6] + dkk
            dk[5] = dk[5] - dkk
        dd = dden * d / (n * den)
        du = du + dd
        dg6 = dg6 - dd
    dps = np.zeros(model.nparams, dtype=dtype)
    # k7 = f(u, t+dt)
    a, dp = model.vjp(u, ps, t + dt, dk[6])
    dps += dp
    du = du + a
    # u = uprev + dt*sum a7i k_i
    for i in range(1, 6):
        dk[i] = dk[i] + dt * tab.a[6][i] * du
    # stages 6..2
    cs = [None, tab.c1, tab.c2, tab.c3, tab.c4, T(1)]
    for j in range(5, 0, -1):        # k_{j+1} = f(g_{j+1}, t + c_j dt)
        g = uprev + dt * _lincomb(tab.a[j], ks)
        a, dp = model.vjp(g, ps, t + cs[j] * dt, dk[j])
        dps += dp
        if j == 5:
            a = a + dg6
        for i in range(1, j):
            dk[i] = dk[i] + dt * tab.a[j][i] * a
    return dps


def _lincomb(coeffs, ks):
    acc = coeffs[0] * ks[0]
    for c, k in zip(coeffs[1:], ks[1:]):
        acc = acc + c * k
    return acc


# --------------------------------------------------------------------------
# Continuous adjoint: InterpolatingAdjoint(autojacvec=ZygoteVJP())
# --------------------------------------------------------------------------
def adjoint_backward(model: MLP, ps, fwd: ODESolution, save_ts, d_us, *, abstol, reltol,
                     maxiters=1000, pow_mode="fastpow_2023"):
    """Integrates z=[vec(lambda); mu] from t2 back to t0 with the same adaptive
    Tsit5.  ``save_ts``/``d_us``: the saved times of the returned solution and the
    cotangent for each (None = zero).  Every saved time is a tstop of the backward
    solve (PresetTimeCallback) where lambda += dL/du(t_i); the jump at t2 is applied
    before fsalfirst/initdt.  Error control is RMS over ALL of z.  Returns
    (d_x, d_ps, backward ODESolution)."""
    dtype = ps.dtype
    T = dtype.type
    t0, t2 = fwd.ts[0], fwd.ts[-1]
    D, B = fwd.us[0].shape
    n = D * B
    P = model.nparams

    def cot(tval):
        acc = None
        for ts_i, d in zip(save_ts, d_us):
            if d is not None and T(ts_i) == T(tval):
                acc = d if acc is None else acc + d
        return acc

    def jump(tval, z):
        d = cot(tval)
        if d is None:
            # the callback still fires and marks u modified
            return z.copy()
        z = z.copy()
        z[:n] += np.asarray(d, dtype=dtype).ravel(order="F")
        return z

    def rhs(z, tval):
        lam = z[:n].reshape((D, B), order="F")
        y = fwd(tval)
        a, dp = model.vjp(y, ps, tval, lam)
        out = np.empty_like(z)
        out[:n] = -a.ravel(order="F")
        out[n:] = -dp
        return out

    z0 = np.zeros(n + P, dtype=dtype)
    stops = [T(s) for s in save_ts if T(s) != t2 and T(s) != t0]
    has_t2 = any(T(s) == t2 for s in save_ts)
    bsol = solve_tsit5(rhs, z0, t2, t0, abstol=abstol, reltol=reltol, maxiters=maxiters,
                       tstops=stops, on_tstop=jump, pow_mode=pow_mode,
                       init_jump=jump if has_t2 else None)
    zf = bsol.us[-1]
    d_x = zf[:n].reshape((D, B), order="F").copy()
    d0 = cot(t0)
    if d0 is not None:
        d_x = d_x + d0
    return d_x, zf[n:].copy(), bsol


# --------------------------------------------------------------------------
# The NeuralODE layer (src/layers/neural_ode.jl)
# --------------------------------------------------------------------------
@dataclass
class LayerOutput:
    """What the functor returns as ``sol``: saved times/states (``sol.t``/``sol.u``)."""
    t: list
    u: list


def diffeqsol_to_array(sol):          # utils.jl:37
    return sol.u[-1]


def diffeqsol_to_timeseries(sol):     # utils.jl:43-45: stack saves on dim ndims-1
    return np.stack(sol.u, axis=1)    # (D, nsave, B)


class NeuralODE:
    """Mirror of ``NeuralODE(model; solver=Tsit5(), sensealg=InterpolatingAdjoint(...),
    tspan=(0f0,1f0), regularize=true, maxiters=1000, regularize_type=:error_estimate,
    kwargs...)`` (neural_ode.jl:10-22).  ``kwargs``: abstol, reltol, saveat, save_start
    (splatted into solve/init at :51,:36; OrdinaryDiffEq defaults abstol=1e-6,
    reltol=1e-3)."""

    VALID_MODES = ("none", "unbiased", "biased")
    VALID_TYPES = ("error_estimate", "stiffness_estimate")

    def __init__(self, model: MLP, *, tspan=(0.0, 1.0), regularize=True, maxiters=1000,
                 regularize_type="error_estimate", dtype=np.float32, pow_mode="fastpow_2023",
                 **kwargs):
        if isinstance(regularize, bool):
            regularize = "unbiased" if regularize else "none"      # :14-16
        if regularize not in self.VALID_MODES:                      # utils.jl:53-58
            raise ValueError(f"regularize must be one of {self.VALID_MODES}")
        if regularize_type not in self.VALID_TYPES:
            raise ValueError(f"regularize must be one of {self.VALID_TYPES}")
        self.model, self.tspan, self.regularize = model, tspan, regularize
        self.regularize_type, self.maxiters = regularize_type, maxiters
        self.dtype = np.dtype(dtype)
        self.pow_mode = pow_mode
        self.abstol = kwargs.pop("abstol", 1e-6)
        self.reltol = kwargs.pop("reltol", 1e-3)
        self.saveat = kwargs.pop("saveat", None)
        self.save_start = kwargs.pop("save_start", None)
        if kwargs:
            raise TypeError(f"unsupported solve kwargs {sorted(kwargs)}")

    def initialstates(self, rng: np.random.Generator):          # neural_ode.jl:27-31
        rng.standard_normal()
        return dict(model={}, nfe=-1, reg_val=self.dtype.type(0), rng=copy.deepcopy(rng),
                    training=True)

    # ---- helpers
    def _f(self, ps):
        return lambda u, t: self.model.f(u, ps, t)

    def _solve(self, x, ps):
        T = self.dtype.type
        return solve_tsit5(self._f(ps), x, T(self.tspan[0]), T(self.tspan[1]),
                           abstol=self.abstol, reltol=self.reltol, maxiters=self.maxiters,
                           pow_mode=self.pow_mode)

    def _saves(self, sol: ODESolution, saveat, all_steps=False):
        T = self.dtype.type
        t0 = T(self.tspan[0])
        if all_steps:     # saveat=[]: every accepted step (+ start unless save_start=false)
            start = 0 if (self.save_start is None or self.save_start) else 1
            return LayerOutput(list(sol.ts[start:]), list(sol.us[start:]))
        times = [T(s) for s in saveat]
        if self.save_start and t0 not in times:
            times = [t0] + times
        # failed solves (retcode != Success) return what exists: clamp to the last time
        tl = sol.ts[-1]
        return LayerOutput(times, [sol(min(s, tl) if sol.tdir > 0 else s,
                                       exact_hit_stored=True) for s in times])

    def __call__(self, x, ps, st):
        """(n::NeuralODE)(x, ps, st) -> (sol, st') (neural_ode.jl:62-100)."""
        out, st2, _ = self.forward(x, ps, st)
        return out, st2

    def forward(self, x, ps, st):
        T = self.dtype.type
        x = np.asarray(x, dtype=self.dtype)
        ps = np.asarray(ps, dtype=self.dtype)
        t0, t2 = T(self.tspan[0]), T(self.tspan[1])
        mode = self.regularize if st["training"] else "none"   # :66, :86
        sol = self._solve(x, ps)
        if mode == "none":                                       # :56-60, :102-105
            if self.saveat is None:
                out = self._saves(sol, [t2])
            else:
                out = self._saves(sol, self.saveat)
            st2 = dict(model=st["model"], nfe=sol.nf, reg_val=T(0), rng=st["rng"],
                       training=st["training"])
            return out, st2, dict(sol=sol, mode=mode, save_ts=list(out.t), x=x, ps=ps)
        rng = copy.deepcopy(st["rng"])                            # Lux.replicate, :69/:91
        if mode == "unbiased":
            t1 = T(rng.random(dtype=np.float32)) * (t2 - t0) + t0   # :71
            if self.saveat is None:
                out_full = self._saves(sol, [t1, t2])             # :108
                out = out_full
            else:
                out_full = self._saves(sol, list(self.saveat) + [t1])   # :109
                keep = [i for i, s in enumerate(out_full.t) if s != t1]  # utils.jl:31-33
                out = LayerOutput([out_full.t[i] for i in keep], [out_full.u[i] for i in keep])
        else:  # biased :88-100
            if self.saveat is None:
                out_full = self._saves(sol, None, all_steps=True)  # :113-116
            else:
                out_full = self._saves(sol, self.saveat)
            out = out_full
            cand = out_full.t[:-1]                                # sol.t[1:end-1], :92
            t1 = cand[int(rng.integers(0, len(cand)))] if cand else t0
        u1 = sol(t1, exact_hit_stored=True)                       # :34
        f = self._f(ps)
        k1 = f(u1, t1)                                            # init: fsalfirst
        dt, _ = ode_initdt(f, u1, t1, 1, abs(t2 - t1), self.abstol, self.reltol,
                           dtmin=max(np.spacing(t1), np.spacing(t2)))
        _, reg, nf2, _, _ = perform_step_tsit5_reg(f, u1, k1, t1, dt, self.abstol,
                                                   self.reltol, self.regularize_type)
        st2 = dict(model=st["model"], nfe=sol.nf + nf2, reg_val=T(reg), rng=rng,
                   training=st["training"])                       # :79-83
        aux = dict(sol=sol, mode=mode, t1=t1, u1=u1, k1=k1, dt_reg=dt, out_full=out_full,
                   x=x, ps=ps)
        return out, st2, aux

    def backward(self, aux, d_us, d_reg, ps):
        """rrule of the functor: cotangents on each returned ``sol.u[i]`` (None = zero)
        and on ``st'.reg_val``.  Returns (d_x, d_ps).  d reg/d x == 0 (runtests.jl:129)."""
        ps = np.asarray(ps, dtype=self.dtype)
        sol = aux["sol"]
        out_full = aux.get("out_full")
        if aux["mode"] == "none":
            save_ts = aux["save_ts"]
            d_full = list(d_us)
        else:
            save_ts = list(out_full.t)
            if aux["mode"] == "unbiased" and self.saveat is not None:
                it = iter(d_us)
                d_full = [None if s == aux["t1"] else next(it) for s in save_ts]
            else:
                d_full = list(d_us)
        d_x, d_ps, bsol = adjoint_backward(self.model, ps, sol, save_ts, d_full,
                                           abstol=self.abstol, reltol=self.reltol,
                                           maxiters=self.maxiters, pow_mode=self.pow_mode)
        if aux["mode"] != "none" and d_reg is not None and d_reg != 0:
            d_ps = d_ps + reg_step_backward(self.model, ps, aux["u1"], aux["k1"], aux["t1"],
                                            aux["dt_reg"], self.abstol, self.reltol,
                                            self.regularize_type, d_reg)
        aux["bsol"] = bsol
        return d_x, d_ps


# --------------------------------------------------------------------------
# SOSRI local-reg step (src/perform_step.jl:49-106); coefficients SURVEY A.6
# --------------------------------------------------------------------------
SOSRI = dict(
    a021=-0.04199224421316468, a031=2.842612915017106, a032=-2.0527723684000727,
    a041=4.338237071435815, a042=-2.8895936137439793, a043=2.3017575594644466,
    a121=0.26204282091330466, a131=0.20903646383505375, a132=-0.1502377115150361,
    a141=0.05836595312746999, a142=0.6149440396332373, a143=0.08535117634046772,
    b021=-0.21641093549612528, b031=1.5336352863679572, b032=0.26066223492647056,
    b041=-1.0536037558179159, b042=1.7015284721089472, b043=-0.20725685784180017,
    b121=-0.5119011827621657, b131=2.67767339866713, b132=-4.9395031322250995,
    b141=0.15580956238299215, b142=3.2361551006624674, b143=-1.4223118283355949,
    alpha1=1.140099274172029, alpha2=-0.6401334255743456, alpha3=0.4736296532772559,
    alpha4=0.026404498125060714,
    c02=-0.04199224421316468, c03=0.7898405466170333, c04=3.7504010171562823,
    c11=0.0, c12=0.26204282091330466, c13=0.05879875232001766, c14=0.758661169101175,
    beta11=-1.8453464565104432, beta12=2.688764531100726, beta13=-0.2523866501071323,
    beta14=0.40896857551684956,
    beta21=0.4969658141589478, beta22=-0.5771202869753592, beta23=-0.12919702470322217,
    beta24=0.2093514975196336,
    beta31=2.8453464565104425, beta32=-2.688764531100725, beta33=0.2523866501071322,
    beta34=-0.40896857551684945,
    beta41=0.11522663875443433, beta42=-0.57877086147738, beta43=0.2857851028163886,
    beta44=0.17775911990655704,
)


def perform_step_sosri_reg(fd, gd, uprev, t, dt, dW, dZ, abstol, reltol, delta=1.0 / 6.0):
    """``_perform_step(integrator, ::FourStageSRIConstantCache, p)`` for diagonal
    noise (perform_step.jl:49-106).  ``fd(u,t)`` drift, ``gd(u,t)`` diffusion.
    Returns (u, EEst*dt, 0, dt)."""
    T = uprev.dtype.type
    c = {k: T(v) for k, v in SOSRI.items()}
    t, dt = T(t), T(dt)
    sqdt = np.sqrt(abs(dt))
    sqrt3 = np.sqrt(T(3))
    chi1 = (dW ** 2 - abs(dt)) / (T(2) * sqdt)
    chi2 = (dW + dZ / sqrt3) / T(2)
    chi3 = (dW ** 3 - T(3) * dW * dt) / (T(6) * dt)
    k1 = fd(uprev, t)
    g1 = gd(uprev, t + c["c11"] * dt)
    H01 = uprev + dt * c["a021"] * k1 + c["b021"] * chi2 * g1
    H11 = uprev + dt * c["a121"] * k1 + sqdt * c["b121"] * g1
    k2 = fd(H01, t + c["c02"] * dt)
    g2 = gd(H11, t + c["c12"] * dt)
    H02 = uprev + dt * (c["a031"] * k1 + c["a032"] * k2) + chi2 * (c["b031"] * g1 + c["b032"] * g2)
    H12 = uprev + dt * (c["a131"] * k1 + c["a132"] * k2) + sqdt * (c["b131"] * g1 + c["b132"] * g2)
    k3 = fd(H02, t + c["c03"] * dt)
    g3 = gd(H12, t + c["c13"] * dt)
    H03 = (uprev + dt * (c["a041"] * k1 + c["a042"] * k2 + c["a043"] * k3)
           + chi2 * (c["b041"] * g1 + c["b042"] * g2 + c["b043"] * g3))
    H13 = (uprev + dt * (c["a141"] * k1 + c["a142"] * k2 + c["a143"] * k3)
           + sqdt * (c["b141"] * g1 + c["b142"] * g2 + c["b143"] * g3))
    k4 = fd(H03, t + c["c04"] * dt)
    g4 = gd(H13, t + c["c14"] * dt)
    E2 = (chi2 * (c["beta31"] * g1 + c["beta32"] * g2 + c["beta33"] * g3 + c["beta34"] * g4)
          + chi3 * (c["beta41"] * g1 + c["beta42"] * g2 + c["beta43"] * g3 + c["beta44"] * g4))
    u = (uprev + dt * (c["alpha1"] * k1 + c["alpha2"] * k2 + c["alpha3"] * k3 + c["alpha4"] * k4)
         + E2
         + dW * (c["beta11"] * g1 + c["beta12"] * g2 + c["beta13"] * g3 + c["beta14"] * g4)
         + chi1 * (c["beta21"] * g1 + c["beta22"] * g2 + c["beta23"] * g3 + c["beta24"] * g4))
    E1 = dt * (k1 + k2 + k3 + k4)
    resid = (T(delta) * E1 + E2) / (T(abstol) + np.maximum(np.abs(uprev), np.abs(u)) * T(reltol))
    EEst = np.sqrt(np.sum(resid * resid, dtype=uprev.dtype) / T(u.size))
    return u, EEst * dt, 0, dt
