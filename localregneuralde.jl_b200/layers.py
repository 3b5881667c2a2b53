"""Host-side mirror of the reference's Lux layer interface for the hot path.

    layer(x, ps, st) -> (sol, st')          src/layers/neural_ode.jl:62

with ``st`` carrying ``nfe`` / ``reg_val`` / ``rng`` / ``training`` exactly as the reference
does (neural_ode.jl:27-31, :83).  All arithmetic happens in libLRNDE.so on the GPU; this file
only resolves keyword arguments, samples the host-side scalar ``t1`` (neural_ode.jl:71) and
passes pointers.  Inputs may be numpy arrays (host buffers, staged inside the library) or torch
CUDA tensors (device pointers, zero-copy).
"""
from __future__ import annotations

import copy
import ctypes as C
import math
import warnings
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import LayerDesc, Opts, Stats, check, lib

try:  # torch is plumbing only (device memory / streams)
    import torch
except Exception:  # pragma: no cover
    torch = None


# ------------------------------------------------------------------ dynamics descriptors
@dataclass(frozen=True)
class Dense:
    """Lux ``Dense(in => out, act)``."""
    in_dims: int
    out_dims: int
    activation: str = "identity"


class Chain:
    """Lux ``Chain(Dense...)`` (parameters named layer_1..layer_N, flattened in that order)."""

    def __init__(self, *layers: Dense, input_activation: Optional[str] = None):
        self.layers: List[Dense] = list(layers)
        self.input_activation = input_activation  # Base.Fix1(broadcast, act) (construct.jl:235)
        self.time_dependent = False


class TDChain(Chain):
    """``TDChain(chain)``: the scalar t is appended as an extra input row before EVERY layer
    (src/layers/common.jl:19-33)."""

    def __init__(self, chain: Chain):
        super().__init__(*chain.layers, input_activation=chain.input_activation)
        self.time_dependent = True


@dataclass(frozen=True)
class Conv:
    """Lux ``Conv((3,3), in_ch => out_ch; pad=(1,1), use_bias=false)``, optionally followed by
    ``BatchNorm(out_ch, activation)`` as ``Chain(Conv, BatchNorm)`` (experiments/src/construct.jl:213-218).
    ``in_ch`` excludes the TDChain time channel."""
    in_ch: int
    out_ch: int
    batchnorm: bool = False
    activation: str = "identity"


class ConvChain:
    """The conv dynamics of the cifar10 config on a WHCN state ``(width, height, channels, B)``, handed to the
    layer as the flat column-major ``(width*height*channels, B)`` array (a reshape, no copy).  Parameters in
    ComponentArray order: ``weight[3,3,in(+1),out]`` then BatchNorm ``scale[out]``, ``bias[out]`` per layer."""

    def __init__(self, *layers: Conv, width: int, height: int):
        self.layers: List[Conv] = list(layers)
        self.width, self.height = int(width), int(height)
        self.time_dependent = False
        self.input_activation = None

    @property
    def state_dims(self) -> int:
        return self.width * self.height * self.layers[0].in_ch


class TDConvChain(ConvChain):
    """``TDChain(Chain(...))`` over conv layers: ``t .* ones`` is concatenated as one more input CHANNEL before
    every outer layer (src/layers/common.jl:19-33)."""

    def __init__(self, chain: ConvChain):
        super().__init__(*chain.layers, width=chain.width, height=chain.height)
        self.time_dependent = True


def initial_model_state(model) -> dict:
    """``Lux.initialstates`` of the dynamics: BatchNorm running_mean = 0 / running_var = 1 per BatchNorm layer
    (flat: mean[C] then var[C], layer order) and ``training`` (``Lux.testmode`` flips it); {} for Dense chains."""
    if not isinstance(model, ConvChain) or not any(L.batchnorm for L in model.layers):
        return {}
    blocks = []
    for L in model.layers:
        if L.batchnorm:
            blocks += [np.zeros(L.out_ch, np.float32), np.ones(L.out_ch, np.float32)]
    return dict(running=np.concatenate(blocks), training=True)


def _state_dims(model) -> int:
    return model.state_dims if isinstance(model, ConvChain) else model.layers[0].in_dims


def nparams(model) -> int:
    td = 1 if model.time_dependent else 0
    if isinstance(model, ConvChain):
        return sum(9 * (L.in_ch + td) * L.out_ch + (2 * L.out_ch if L.batchnorm else 0) for L in model.layers)
    return sum(L.out_dims * (L.in_dims + td) + L.out_dims for L in model.layers)


def glorot_uniform(model, rng: np.random.Generator) -> np.ndarray:
    """Lux default init (Glorot-uniform weights, zero bias; BatchNorm scale 1, bias 0), flat ComponentArray order."""
    td = 1 if model.time_dependent else 0
    out = []
    if isinstance(model, ConvChain):
        for L in model.layers:
            cin = L.in_ch + td
            a = math.sqrt(6.0 / (9 * cin + 9 * L.out_ch))
            out.append(rng.uniform(-a, a, size=9 * cin * L.out_ch).astype(np.float32))
            if L.batchnorm:
                out.append(np.ones(L.out_ch, np.float32))
                out.append(np.zeros(L.out_ch, np.float32))
        return np.concatenate(out)
    for L in model.layers:
        fan_in, fan_out = L.in_dims + td, L.out_dims
        a = math.sqrt(6.0 / (fan_in + fan_out))
        W = rng.uniform(-a, a, size=(fan_out, fan_in)).astype(np.float32)
        out.append(W.ravel(order="F"))
        out.append(np.zeros(fan_out, np.float32))
    return np.concatenate(out)


# ------------------------------------------------------------------ context
class Context:
    """One device + one stream (include/lrnde.h: a ctx is not thread-safe)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = C.c_void_p()
        check(lib().lrnde_ctx_create(C.byref(self._h), int(device), C.c_void_p(stream or 0)))
        self.device = device
        self._models = {}

    def sync(self):
        check(lib().lrnde_ctx_sync(self._h))

    def set_tape_budget(self, nbytes: int):
        check(lib().lrnde_ctx_set_tape_budget(self._h, int(nbytes)))

    def close(self):
        if self._h:
            for _m, h in self._models.values():
                lib().lrnde_model_destroy(h)
            self._models.clear()
            lib().lrnde_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def model_handle(self, model: Chain):
        key = id(model)
        if key in self._models and self._models[key][0] is not model:
            lib().lrnde_model_destroy(self._models.pop(key)[1])    # id() reuse of a dead object
        if key not in self._models and isinstance(model, ConvChain):
            arr = (_lib.ConvLayerDesc * len(model.layers))()
            for i, L in enumerate(model.layers):
                arr[i] = _lib.ConvLayerDesc(L.in_ch, L.out_ch, 1 if L.batchnorm else 0, _lib.ACT[L.activation])
            h = C.c_void_p()
            check(lib().lrnde_conv_model_create(self._h, arr, len(model.layers), model.width, model.height,
                                                1 if model.time_dependent else 0, C.byref(h)))
            self._models[key] = (model, h)
        if key not in self._models:
            arr = (LayerDesc * len(model.layers))()
            for i, L in enumerate(model.layers):
                arr[i] = LayerDesc(L.in_dims, L.out_dims, _lib.ACT[L.activation])
            h = C.c_void_p()
            check(lib().lrnde_model_create(self._h, arr, len(model.layers),
                                           1 if model.time_dependent else 0,
                                           _lib.ACT[model.input_activation], C.byref(h)))
            self._models[key] = (model, h)       # the strong reference keeps id(model) unique
        return self._models[key][1]

    # ---- data-parallel group (one process per GPU): mailboxes exchanged through CUDA IPC
    def setup_group(self, rank: int, nranks: int, total_batch: int, allgather_bytes):
        """``allgather_bytes(b: bytes) -> list[bytes]`` gathers one blob per rank (e.g. via
        torch.distributed.all_gather_object)."""
        if nranks == 1:
            check(lib().lrnde_ctx_set_dist(self._h, 0, 1, None, int(total_batch)))
            return
        ptr = C.c_void_p()
        nb = C.c_uint64()
        check(lib().lrnde_ctx_mailbox(self._h, C.byref(ptr), C.byref(nb)))
        handle = (C.c_ubyte * 64)()
        check(lib().lrnde_ipc_export(self._h, ptr, handle))
        blobs = allgather_bytes(bytes(handle))
        boxes = (C.c_void_p * nranks)()
        for r, blob in enumerate(blobs):
            if r == rank:
                boxes[r] = ptr
            else:
                buf = (C.c_ubyte * 64).from_buffer_copy(blob)
                p = C.c_void_p()
                check(lib().lrnde_ipc_open(self._h, buf, C.byref(p)))
                boxes[r] = p
        check(lib().lrnde_ctx_set_dist(self._h, rank, nranks, boxes, int(total_batch)))


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        stream = None
        if torch is not None and torch.cuda.is_available():
            stream = torch.cuda.current_stream(device).cuda_stream
        _default_ctx[device] = Context(device, stream)
    return _default_ctx[device]


# ------------------------------------------------------------------ array plumbing
def _is_torch(a):
    return torch is not None and isinstance(a, torch.Tensor)


def _ptr(a):
    if a is None:
        return C.c_void_p(0)
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _host_empty(shape):
    """Host output buffer of the host-pointer path: page-locked when torch can provide it (its caching
    host allocator recycles the blocks), so that the library's device->host copies run at full
    PCIe rate instead of through the driver's pageable staging."""
    if torch is not None and torch.cuda.is_available() and 4 * int(np.prod(shape)) <= (1 << 30):
        try:
            return torch.empty(tuple(shape), dtype=torch.float32, pin_memory=True).numpy()
        except Exception:  # pragma: no cover
            pass
    return np.empty(shape, np.float32)


def _as_input(a, like_torch: bool, device):
    """Column-major (features, batch) -> the flat buffer the C ABI takes."""
    if _is_torch(a):
        if not a.is_cuda:
            raise ValueError("torch inputs must be CUDA tensors (or pass numpy arrays)")
        # (D, B) logical, stored batch-major == column-major flat buffer
        return a.detach().to(torch.float32).t().contiguous()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).T)


class DESolution:
    """What the functor returns as ``sol``: ``sol.u`` (list of (D,B) arrays at ``sol.t``)."""

    def __init__(self, t, u, tape, layer, ctx, stats, host):
        self.t, self.u = t, u
        self._tape, self._layer, self._ctx, self.stats, self._host = tape, layer, ctx, stats, host
        self.retcode = _lib.RETCODES.get(stats.retcode, "?")

    def step_log(self, which: int = 0):
        """(t, dt, EEst, accepted) of every attempted step of the forward (0) / adjoint (1)."""
        if not self._tape:
            raise _lib.LrndeError(-4, "no tape (forward ran with keep_tape=False)")
        n = C.c_int32()
        check(lib().lrnde_step_log(self._tape, which, None, None, None, None, 0, C.byref(n)))
        k = n.value
        t = np.zeros(k, np.float32); dt = np.zeros(k, np.float32)
        e = np.zeros(k, np.float32); a = np.zeros(k, np.uint8)
        check(lib().lrnde_step_log(self._tape, which, _ptr(t), _ptr(dt), _ptr(e), _ptr(a), k,
                                   C.byref(n)))
        return t, dt, e, a.astype(bool)

    def free(self):
        if self._tape:
            lib().lrnde_tape_free(self._tape)
            self._tape = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def diffeqsol_to_array(sol):                      # src/utils.jl:37
    return sol.u[-1]


def diffeqsol_to_timeseries(sol):                 # src/utils.jl:43-45
    if _is_torch(sol.u[0]):
        return torch.stack(sol.u, dim=1)
    return np.stack(sol.u, axis=1)


# ------------------------------------------------------------------ the layer
class NeuralODE:
    """``NeuralODE(model; solver=Tsit5(), sensealg=InterpolatingAdjoint(autojacvec=ZygoteVJP()),
    tspan=(0f0,1f0), regularize=true, maxiters=1000, regularize_type=:error_estimate,
    kwargs...)`` (src/layers/neural_ode.jl:10-22).  kwargs: abstol, reltol, saveat, save_start
    (splatted into solve/init, :51/:36).  Extra, GPU-side only: ``precision`` ("auto", "fp32",
    "tf32x3", "tf32"), ``pow_mode``, ``loop_mode``."""

    VALID_MODES = ("none", "unbiased", "biased")
    VALID_TYPES = ("error_estimate", "stiffness_estimate")

    def __init__(self, model: Chain, *, solver="Tsit5", sensealg="InterpolatingAdjoint",
                 tspan=(0.0, 1.0), regularize=True, maxiters: int = 1000,
                 regularize_type: str = "error_estimate", precision: str = "auto",
                 pow_mode: str = "fastpow_2023", loop_mode: int = 0, ctx: Optional[Context] = None,
                 return_last_only: bool = False, **kwargs):
        if isinstance(regularize, bool):
            regularize = "unbiased" if regularize else "none"          # :14-16
        if regularize not in self.VALID_MODES:                          # utils.jl:53-58
            raise ValueError(f"regularize must be one of {self.VALID_MODES}")
        if regularize_type not in self.VALID_TYPES:
            raise ValueError(f"regularize must be one of {self.VALID_TYPES}")
        if solver != "Tsit5":
            raise ValueError("the B200 path implements solver=Tsit5() (the reference default)")
        if sensealg != "InterpolatingAdjoint":
            raise ValueError("the B200 path implements sensealg=InterpolatingAdjoint(ZygoteVJP)")
        self.model, self.tspan, self.regularize = model, tuple(tspan), regularize
        self.regularize_type, self.maxiters = regularize_type, int(maxiters)
        self.abstol = float(kwargs.pop("abstol", 1e-6))     # OrdinaryDiffEq defaults
        self.reltol = float(kwargs.pop("reltol", 1e-3))
        self.saveat = kwargs.pop("saveat", None)
        self.save_start = kwargs.pop("save_start", None)
        if kwargs:
            raise TypeError(f"unsupported solve kwargs {sorted(kwargs)}")
        self.precision, self.pow_mode, self.loop_mode = precision, pow_mode, loop_mode
        self.return_last_only = bool(return_last_only)   # fused diffeqsol_to_array: sol.u == [u(t2)]
        self._ctx = ctx

    # Lux.initialparameters / initialstates (neural_ode.jl:27-31)
    def initialparameters(self, rng: np.random.Generator) -> np.ndarray:
        return glorot_uniform(self.model, rng)

    def initialstates(self, rng: np.random.Generator):
        rng.standard_normal()
        return dict(model=initial_model_state(self.model), nfe=-1, reg_val=np.float32(0), rng=copy.deepcopy(rng),
                    training=True)

    def _attach_model_state(self, o, mstate, host, device=None):
        """st.model -> lrnde_opts.model_state (a fresh copy that the call updates = st'.model)."""
        if not mstate or mstate.get("running") is None:
            return None
        r = mstate["running"]
        if host:
            run = np.ascontiguousarray(np.asarray(r.detach().cpu() if _is_torch(r) else r, dtype=np.float32)).copy()
        else:
            run = (r.detach().to(device=device, dtype=torch.float32).clone() if _is_torch(r)
                   else torch.from_numpy(np.asarray(r, np.float32).copy()).to(device))
        o.model_state = _ptr(run)
        o.model_testmode = 0 if mstate.get("training", True) else 1
        return run

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context(0)
        return self._ctx

    def _opts(self, mode, t1, u01, keep_tape, host):
        o = Opts()
        o.t0, o.t2 = self.tspan
        o.abstol, o.reltol, o.maxiters = self.abstol, self.reltol, self.maxiters
        o.reg_mode = _lib.REG[mode]
        o.reg_type = _lib.REGTYPE[self.regularize_type]
        o.t1, o.u01 = float(t1), float(u01)
        keep = None
        if self.saveat is not None:
            keep = np.asarray(self.saveat, dtype=np.float32)
            o.saveat = keep.ctypes.data_as(C.POINTER(C.c_float))
            o.nsave = keep.size
        elif mode == "biased":
            o.nsave = -1
        else:
            o.nsave = 0
        # OrdinaryDiffEq: save_start defaults to true only for the save-everything case
        if self.save_start is None:
            o.save_start = 1 if (self.saveat is None and mode == "biased") else 0
        else:
            o.save_start = 1 if self.save_start else 0
        o.precision = _lib.PREC[self.precision]
        o.pow_mode = _lib.POW[self.pow_mode]
        o.host_buffers = 1 if host else 0
        o.keep_tape = 1 if keep_tape else 0
        o.loop_mode = int(self.loop_mode)
        o.last_only = 1 if self.return_last_only else 0
        return o, keep

    def __call__(self, x, ps, st, keep_tape: Optional[bool] = None):
        """(n::NeuralODE)(x, ps, st) -> (sol, st')."""
        T = np.float32
        t0, t2 = T(self.tspan[0]), T(self.tspan[1])
        training = bool(st["training"])
        mode = self.regularize if training else "none"                  # :66, :86
        rng = st["rng"]
        t1, u01 = t0, 0.0
        if mode != "none":
            rng = copy.deepcopy(st["rng"])                              # Lux.replicate (:69)
            if mode == "unbiased":
                t1 = T(rng.random(dtype=np.float32)) * (t2 - t0) + t0   # :71
            else:
                u01 = float(rng.random(dtype=np.float32))               # :92 (index sampled in C)
        host = not _is_torch(x)
        if keep_tape is None:
            keep_tape = training
        o, _keep = self._opts(mode, t1, u01, keep_tape, host)
        xb = _as_input(x, not host, None)
        B, D = xb.shape
        if D != _state_dims(self.model):
            raise ValueError(f"x has {D} features, the dynamics expect {_state_dims(self.model)}")
        psb = ps.detach().to(torch.float32).contiguous() if _is_torch(ps) else \
            np.ascontiguousarray(np.asarray(ps, dtype=np.float32))
        if (not host) != _is_torch(psb):
            raise ValueError("x and ps must both be numpy arrays or both be CUDA tensors")
        n_ps = psb.size if host else psb.numel()
        if n_ps != nparams(self.model):
            raise ValueError("ps has the wrong length for this model")
        def _alloc(cap):
            if host:
                return _host_empty((cap, B, D))
            return torch.empty((cap, B, D), dtype=torch.float32, device=xb.device)

        stats = Stats()
        tape = C.c_void_p()
        ctx = self.ctx
        mh = ctx.model_handle(self.model)
        running = self._attach_model_state(o, st.get("model"), host, None if host else xb.device)
        if o.nsave >= 0:
            cap = (o.nsave + 1) if o.nsave > 0 else 3        # [u0,] [u(t1),] u(t2)
            usave = _alloc(cap)
            times = np.zeros(cap, np.float32)
            check(lib().lrnde_ode_forward(ctx._h, mh, C.byref(o), _ptr(psb), _ptr(xb), B, _ptr(usave),
                                          cap, _ptr(times), C.byref(stats), C.byref(tape)))
        else:
            # every accepted step (:biased, neural_ode.jl:86-100): their number is only known after the solve, so the
            # states are fetched from the kept solution afterwards instead of reserving maxiters + 2 blocks up front
            user_keep = bool(o.keep_tape)
            o.keep_tape = 1
            check(lib().lrnde_ode_forward(ctx._h, mh, C.byref(o), _ptr(psb), _ptr(xb), B, None,
                                          0, None, C.byref(stats), C.byref(tape)))
            cap = max(1, stats.nsave_out)
            usave = _alloc(cap)
            times = np.zeros(cap, np.float32)
            check(lib().lrnde_ode_saved_states(ctx._h, mh, tape, _ptr(usave), cap, _ptr(times)))
            if not user_keep:
                check(lib().lrnde_tape_free(tape))
                tape = C.c_void_p()
        if stats.retcode != 0:
            warnings.warn(f"NeuralODE solve finished with retcode {_lib.RETCODES.get(stats.retcode, stats.retcode)} "
                          "(the reference warns and returns the partial solution as well)")
        n = stats.nsave_out
        us = [usave[i].T for i in range(n)]        # back to (D, B) views
        sol = DESolution([T(t) for t in times[:n]], us, tape if keep_tape else None, self, ctx,
                         stats, host)
        sol._shape = (B, D)
        model_st = st["model"] if running is None else dict(st["model"], running=running)   # st_ of the closure
        st2 = dict(model=model_st, nfe=int(stats.nfe), reg_val=T(stats.reg_val), rng=rng,
                   training=st["training"])                             # :79-83
        return sol, st2

    def backward(self, sol: DESolution, d_us: Sequence, d_reg: float = 0.0):
        """Pullback of the functor: cotangents on each ``sol.u[i]`` (None = zero) and on
        ``st'.reg_val``.  Returns (d_x, d_ps); d reg / d x == 0 (test/runtests.jl:129)."""
        if not sol._tape:
            raise _lib.LrndeError(-4, "no tape: call the layer with training=True / keep_tape=True")
        B, D = sol._shape
        n = len(sol.u)
        if len(d_us) != n:
            raise ValueError(f"expected {n} cotangent blocks, got {len(d_us)}")
        host = sol._host
        P = nparams(self.model)
        if all(d is None for d in d_us):
            dU = None
        elif host:
            dU = _host_empty((n, B, D))
            for i, d in enumerate(d_us):
                if d is not None:
                    dU[i] = np.asarray(d, np.float32).T
                else:
                    dU[i] = 0.0
        else:
            dev = sol.u[0].device
            dU = torch.zeros((n, B, D), dtype=torch.float32, device=dev)
            for i, d in enumerate(d_us):
                if d is not None:
                    dU[i] = d.to(torch.float32).t()
        if host:
            d_x = _host_empty((B, D))
            d_ps = _host_empty((P,))
        else:
            dev = sol.u[0].device
            d_x = torch.empty((B, D), dtype=torch.float32, device=dev)
            d_ps = torch.empty(P, dtype=torch.float32, device=dev)
        stats = Stats()
        ctx = sol._ctx
        check(lib().lrnde_ode_backward(ctx._h, ctx.model_handle(self.model), sol._tape, _ptr(dU),
                                       float(d_reg), _ptr(d_ps), _ptr(d_x), C.byref(stats)))
        sol.bwd_stats = stats
        return d_x.T, d_ps

    def dynamics(self, u, ps, t: float, model_state: Optional[dict] = None):
        """One evaluation of the ``dudt`` closure (neural_ode.jl:45-48) on the GPU.  ``model_state`` (st.model of
        conv dynamics with BatchNorm): returns ``(du, st_model')`` like ``Lux.apply`` does."""
        host = not _is_torch(u)
        ub = _as_input(u, not host, None)
        B, D = ub.shape
        psb = ps.contiguous() if _is_torch(ps) else np.ascontiguousarray(np.asarray(ps, np.float32))
        out = np.empty((B, D), np.float32) if host else torch.empty_like(ub)
        o, _ = self._opts("none", 0.0, 0.0, False, host)
        ctx = self.ctx
        running = self._attach_model_state(o, model_state, host, None if host else ub.device)
        check(lib().lrnde_dynamics_eval(ctx._h, ctx.model_handle(self.model), C.byref(o), _ptr(psb),
                                        _ptr(ub), float(t), B, _ptr(out)))
        if model_state is not None:
            return out.T, (model_state if running is None else dict(model_state, running=running))
        return out.T

    def dynamics_vjp(self, u, ps, t: float, lam, model_state: Optional[dict] = None):
        """``(J_u^T lam, J_p^T lam)`` of one ``dudt`` evaluation: the ZygoteVJP pullback the adjoint calls
        (neural_ode.jl:11).  Host arrays only (parity hook)."""
        ub = _as_input(u, False, None)
        lb = _as_input(lam, False, None)
        B, D = ub.shape
        psb = np.ascontiguousarray(np.asarray(ps, np.float32))
        a = np.empty((B, D), np.float32)
        dps = np.empty(psb.size, np.float32)
        o, _ = self._opts("none", 0.0, 0.0, False, True)
        ctx = self.ctx
        _running = self._attach_model_state(o, model_state, True)
        check(lib().lrnde_dynamics_vjp(ctx._h, ctx.model_handle(self.model), C.byref(o), _ptr(psb), _ptr(ub),
                                       float(t), _ptr(lb), B, _ptr(a), _ptr(dps)))
        return a.T, dps


# ------------------------------------------------------------------ neural SDE layer
class SDESolution:
    """``sol`` of the NeuralDSDE functor: ``sol.u`` (list of (D,B) arrays at ``sol.t``)."""

    def __init__(self, t, u, tape, layer, ctx, stats, host, shape, dev_times, t1_block):
        self.t, self.u = t, u
        self._tape, self._layer, self._ctx, self.stats, self._host = tape, layer, ctx, stats, host
        self._shape, self._dev_times, self._t1_block = shape, dev_times, t1_block
        self.retcode = _lib.RETCODES.get(stats.retcode, "?")

    def step_log(self):
        if not self._tape:
            raise _lib.LrndeError(-4, "no tape (forward ran with keep_tape=False)")
        n = C.c_int32()
        check(lib().lrnde_sde_step_log(self._tape, None, None, None, None, 0, C.byref(n)))
        k = n.value
        t = np.zeros(k, np.float32); dt = np.zeros(k, np.float32)
        e = np.zeros(k, np.float32); a = np.zeros(k, np.uint8)
        check(lib().lrnde_sde_step_log(self._tape, _ptr(t), _ptr(dt), _ptr(e), _ptr(a), k, C.byref(n)))
        return t, dt, e, a.astype(bool)

    def free(self):
        if self._tape:
            lib().lrnde_sde_tape_free(self._tape)
            self._tape = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class NeuralDSDE:
    """``NeuralDSDE(drift, diffusion; solver=SOSRI(), sensealg=TrackerAdjoint(), tspan=(0f0,1f0),
    regularize=:unbiased, maxiters=1000, kwargs...)`` (src/layers/neural_sde.jl:11-19); kwargs:
    abstol, reltol, saveat, save_start.  Parameters are ``[ps.drift; ps.diffusion]`` flat.
    ``seed`` keys the Philox Wiener process (the reference's is unseeded, :68-69)."""

    VALID_MODES = ("none", "unbiased", "biased")

    def __init__(self, drift: Chain, diffusion: Chain, *, solver="SOSRI", sensealg="TrackerAdjoint",
                 tspan=(0.0, 1.0), regularize: str = "unbiased", maxiters: int = 1000,
                 pow_mode: str = "fastpow_2023", seed: int = 0, ctx: Optional[Context] = None,
                 controller: Optional[dict] = None, **kwargs):
        if regularize not in self.VALID_MODES:                          # utils.jl:53-58
            raise ValueError(f"regularize must be one of {self.VALID_MODES}")
        if solver != "SOSRI":
            raise ValueError("the B200 path implements solver=SOSRI() (the reference default)")
        if sensealg != "TrackerAdjoint":
            raise ValueError("the B200 path implements sensealg=TrackerAdjoint()")
        self.drift, self.diffusion, self.tspan, self.regularize = drift, diffusion, tuple(tspan), regularize
        self.maxiters, self.pow_mode, self.seed = int(maxiters), pow_mode, int(seed)
        self.abstol = float(kwargs.pop("abstol", 1e-2))     # StochasticDiffEq defaults
        self.reltol = float(kwargs.pop("reltol", 1e-2))
        self.saveat = kwargs.pop("saveat", None)
        self.save_start = kwargs.pop("save_start", None)
        if kwargs:
            raise TypeError(f"unsupported solve kwargs {sorted(kwargs)}")
        self.controller = dict(controller or {})
        self._ctx = ctx

    def initialparameters(self, rng: np.random.Generator) -> np.ndarray:
        return np.concatenate([glorot_uniform(self.drift, rng), glorot_uniform(self.diffusion, rng)])

    def initialstates(self, rng: np.random.Generator):                   # neural_sde.jl:22-27
        rng.standard_normal()
        return dict(drift={}, diffusion={}, nfe_drift=-1, nfe_diffusion=-1, reg_val=np.float32(0),
                    rng=copy.deepcopy(rng), training=True)

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context(0)
        return self._ctx

    def __call__(self, x, ps, st, keep_tape: Optional[bool] = None):
        """(n::NeuralDSDE)(x, ps, st) -> (sol, st') (neural_sde.jl:82-123)."""
        T = np.float32
        t0, t2 = T(self.tspan[0]), T(self.tspan[1])
        training = bool(st["training"])
        mode = self.regularize if training else "none"                   # :86, :108
        rng = st["rng"]
        t1, u01 = t0, 0.0
        if mode != "none":
            rng = copy.deepcopy(st["rng"])                               # Lux.replicate (:87)
            if mode == "unbiased":
                t1 = T(rng.random(dtype=np.float32)) * (t2 - t0) + t0    # :89
            else:
                u01 = float(rng.random(dtype=np.float32))                # :112 (index picked in C)
        host = not _is_torch(x)
        if keep_tape is None:
            keep_tape = training
        # ---- the save times the device sees (sorted) and how they map back to sol.u
        user = None if self.saveat is None else [T(s) for s in self.saveat]
        if user is not None and self.save_start and t0 not in user:
            user = [t0] + user
        all_steps = False
        if mode == "unbiased":
            ret_times = [t1, t2] if user is None else list(user)         # :90-91, utils.jl:31-33
            dev_times = sorted(set(ret_times + [t1]))
        elif mode == "biased" and user is None:
            all_steps, ret_times, dev_times = True, None, []
        else:
            ret_times = [t2] if user is None else list(user)
            dev_times = sorted(set(ret_times))
        o = _lib.SdeOpts()
        o.t0, o.t2, o.abstol, o.reltol, o.maxiters = t0, t2, self.abstol, self.reltol, self.maxiters
        o.reg_mode, o.t1, o.u01 = _lib.REG[mode], float(t1), u01
        keep = np.asarray(dev_times, dtype=np.float32)
        if all_steps:
            o.nsave = -1
            o.save_start = 1 if (self.save_start is None or self.save_start) else 0
        else:
            o.saveat = keep.ctypes.data_as(C.POINTER(C.c_float))
            o.nsave = keep.size
        o.seed, o.pow_mode = self.seed, _lib.POW[self.pow_mode]
        o.host_buffers, o.keep_tape = (1 if host else 0), (1 if keep_tape else 0)
        for k in ("qmax", "gamma", "qmin", "delta"):
            setattr(o, k, float(self.controller.get(k, 0.0)))
        xb = _as_input(x, not host, None)
        B, D = xb.shape
        psb = ps.detach().to(torch.float32).contiguous() if _is_torch(ps) else \
            np.ascontiguousarray(np.asarray(ps, dtype=np.float32))
        nf, ng = nparams(self.drift), nparams(self.diffusion)
        if (psb.size if host else psb.numel()) != nf + ng:
            raise ValueError("ps has the wrong length for [drift; diffusion]")
        pf, pg = psb[:nf], psb[nf:]
        nblk = max(1, len(dev_times))
        usave = np.empty((nblk, B, D), np.float32) if host else \
            torch.empty((nblk, B, D), dtype=torch.float32, device=xb.device)
        stats = _lib.SdeStats()
        tape = C.c_void_p()
        ctx = self.ctx
        check(lib().lrnde_sde_forward(ctx._h, ctx.model_handle(self.drift), ctx.model_handle(self.diffusion),
                                      C.byref(o), _ptr(pf), _ptr(pg), _ptr(xb), B, _ptr(usave),
                                      C.byref(stats), C.byref(tape)))
        try:
            if all_steps:
                n = stats.nsave_out
                first = 0 if o.save_start else 1
                states = np.empty((max(n, 1), B, D), np.float32) if host else \
                    torch.empty((max(n, 1), B, D), dtype=torch.float32, device=xb.device)
                times = np.zeros(max(n, 1), np.float32)
                if not tape:
                    raise _lib.LrndeError(-4, "regularize=:biased without saveat needs keep_tape=True")
                check(lib().lrnde_sde_states(ctx._h, tape, first, n, 1 if host else 0, _ptr(states), _ptr(times)))
                ts, us = [T(t) for t in times[:n]], [states[i].T for i in range(n)]
                t1_block = None
            else:
                index = {float(s): i for i, s in enumerate(dev_times)}
                ts = [T(s) for s in ret_times]
                us = [usave[index[float(s)]].T for s in ret_times]
                t1_block = index.get(float(t1)) if (mode == "unbiased" and user is not None) else None
        except Exception:
            if tape:
                lib().lrnde_sde_tape_free(tape)
            raise
        sol = SDESolution(ts, us, tape if keep_tape else None, self, ctx, stats, host, (B, D),
                          [float(s) for s in dev_times], t1_block)
        st2 = dict(st, nfe_drift=int(stats.nfe_drift), nfe_diffusion=int(stats.nfe_diffusion),
                   reg_val=T(stats.reg_val), rng=rng)                     # :103-105
        return sol, st2

    def backward(self, sol: SDESolution, d_us: Sequence, d_reg: float = 0.0):
        """TrackerAdjoint pullback: cotangents on each ``sol.u[i]`` (None = zero) and on
        ``st'.reg_val``.  Returns (d_x, d_ps) with d_ps = [d_drift; d_diffusion]."""
        if not sol._tape:
            raise _lib.LrndeError(-4, "no tape: call the layer with training=True / keep_tape=True")
        B, D = sol._shape
        if len(d_us) != len(sol.u):
            raise ValueError(f"expected {len(sol.u)} cotangent blocks, got {len(d_us)}")
        host = sol._host
        nblk = len(sol._dev_times) if sol._dev_times else len(sol.u)
        dev = None if host else sol.u[0].device
        dU = np.zeros((max(nblk, 1), B, D), np.float32) if host else \
            torch.zeros((max(nblk, 1), B, D), dtype=torch.float32, device=dev)
        index = {s: i for i, s in enumerate(sol._dev_times)}
        for i, d in enumerate(d_us):
            if d is None:
                continue
            k = index[float(sol.t[i])] if sol._dev_times else i
            dU[k] += (np.asarray(d, np.float32).T if host else d.to(torch.float32).t())
        nf, ng = nparams(self.drift), nparams(self.diffusion)
        if host:
            d_x, d_ps = np.empty((B, D), np.float32), np.empty(nf + ng, np.float32)
        else:
            d_x = torch.empty((B, D), dtype=torch.float32, device=dev)
            d_ps = torch.empty(nf + ng, dtype=torch.float32, device=dev)
        ctx = sol._ctx
        check(lib().lrnde_sde_backward(ctx._h, sol._tape, _ptr(dU), float(d_reg), _ptr(d_ps[:nf]),
                                       _ptr(d_ps[nf:]), _ptr(d_x)))
        return d_x.T, d_ps

    def perform_step(self, kind: str, ps, uprev, dW, t: float, dt: float, delta: float = 1.0 / 6.0):
        """``_perform_step`` of the alternative SDE caches (src/perform_step.jl:108-206) with injected
        increments: kind "RKMilCommute" or "LambaEulerHeun".  Returns (u, reg_val = EEst*dt)."""
        kinds = {"RKMilCommute": 0, "LambaEulerHeun": 1}
        if kind not in kinds:
            raise ValueError(f"kind must be one of {sorted(kinds)}")
        host = not _is_torch(uprev)
        ub, wb = _as_input(uprev, not host, None), _as_input(dW, not host, None)
        B, D = ub.shape
        psb = ps.contiguous() if _is_torch(ps) else np.ascontiguousarray(np.asarray(ps, np.float32))
        nf = nparams(self.drift)
        out = np.empty((B, D), np.float32) if host else torch.empty_like(ub)
        reg = C.c_float()
        ctx = self.ctx
        check(lib().lrnde_sde_aux_step(ctx._h, ctx.model_handle(self.drift), ctx.model_handle(self.diffusion),
                                       kinds[kind], _ptr(psb[:nf]), _ptr(psb[nf:]), _ptr(ub), _ptr(wb),
                                       float(t), float(dt), self.abstol, self.reltol, float(delta), B,
                                       1 if host else 0, _ptr(out), C.byref(reg)))
        return out.T, np.float32(reg.value)


# ------------------------------------------------------------------ latent-ODE encoder
class LatentGRUCell:
    """``LatentGRUCell(in_dim, h_dim, latent_dim)`` (src/layers/latent_ode.jl:10-17): three two-layer gate
    networks on ``vcat(y_mean, y_std, x)`` with ``x`` of ``2*in_dim + 1`` rows (data, mask, dt)."""

    def __init__(self, in_dim: int, h_dim: int, latent_dim: int):
        self.in_dim, self.h_dim, self.latent_dim = int(in_dim), int(h_dim), int(latent_dim)
        self.features = 2 * self.in_dim + 1

    def nparams(self) -> int:
        return int(lib().lrnde_gru_nparams(self.features, self.h_dim, self.latent_dim))

    def initialparameters(self, rng: np.random.Generator) -> np.ndarray:
        """Lux defaults (Glorot-uniform weights, zero biases) in ComponentArray order
        update_gate, reset_gate, new_state."""
        L, H, I = self.latent_dim, self.h_dim, 2 * self.latent_dim + self.features
        out = []
        for out2 in (L, L, 2 * L):
            for (i, o) in ((I, H), (H, out2)):
                a = math.sqrt(6.0 / (i + o))
                out.append(rng.uniform(-a, a, size=(o, i)).astype(np.float32).ravel(order="F"))
                out.append(np.zeros(o, np.float32))
        return np.concatenate(out)


class Recurrence:
    """``Recurrence(cell)`` as used at experiments/src/construct.jl:231: the cell runs along the time
    dimension of ``x`` (features, time, batch) and the last output ``vcat(y_mean, y_std)`` is returned.
    One kernel launch walks the whole series; ``backward`` is back-propagation through time w.r.t. ``ps``."""

    def __init__(self, cell: LatentGRUCell, ctx: Optional[Context] = None):
        self.cell, self._ctx = cell, ctx

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context(0)
        return self._ctx

    def __call__(self, x, ps, st=None, keep_tape: bool = True):
        c = self.cell
        host = not _is_torch(x)
        if x.shape[0] != c.features:
            raise ValueError(f"x has {x.shape[0]} feature rows, the cell expects {c.features}")
        F, T, B = x.shape
        if host:
            xb = np.ascontiguousarray(np.transpose(np.asarray(x, np.float32), (2, 1, 0)))     # (B, T, F) C order
            psb = np.ascontiguousarray(np.asarray(ps, np.float32))
            y = _host_empty((B, 2 * c.latent_dim))
        else:
            xb = x.detach().to(torch.float32).permute(2, 1, 0).contiguous()
            psb = ps.detach().to(torch.float32).contiguous()
            y = torch.empty((B, 2 * c.latent_dim), dtype=torch.float32, device=x.device)
        if (psb.size if host else psb.numel()) != c.nparams():
            raise ValueError("ps has the wrong length for this cell")
        tape = C.c_void_p()
        check(lib().lrnde_gru_forward(self.ctx._h, c.features, c.h_dim, c.latent_dim, _ptr(psb), _ptr(xb), T, B,
                                      1 if host else 0, 1 if keep_tape else 0, _ptr(y), C.byref(tape)))
        return y.T, dict(st or {}, tape=tape if keep_tape else None, host=host, B=B)

    def backward(self, st, d_y):
        """d_ps for a cotangent on the returned ``y`` (2*latent, B)."""
        if not st.get("tape"):
            raise _lib.LrndeError(-4, "no tape: call the layer with keep_tape=True")
        host, c = st["host"], self.cell
        if host:
            dy = np.ascontiguousarray(np.asarray(d_y, np.float32).T)
            d_ps = np.empty(c.nparams(), np.float32)
        else:
            dy = d_y.to(torch.float32).t().contiguous()
            d_ps = torch.empty(c.nparams(), dtype=torch.float32, device=d_y.device)
        check(lib().lrnde_gru_backward(self.ctx._h, st["tape"], _ptr(dy), _ptr(d_ps)))
        return d_ps

    def free(self, st):
        if st.get("tape"):
            lib().lrnde_gru_tape_free(st["tape"])
            st["tape"] = None


def _cols(a):
    """Julia column-major (rows, d2, ..., dn) -> the flat buffer the C ABI takes (reversed axes, C order)."""
    if _is_torch(a):
        return a.detach().to(torch.float32).permute(*reversed(range(a.dim()))).contiguous()
    a = np.asarray(a, np.float32)
    return np.ascontiguousarray(a.transpose(tuple(reversed(range(a.ndim)))))


def _out_like(shape_cols, like):
    """Output buffer with C shape ``shape_cols`` (reversed Julia shape) on the side ``like`` lives on."""
    if _is_torch(like):
        return torch.empty(tuple(shape_cols), dtype=torch.float32, device=like.device)
    return np.empty(tuple(shape_cols), np.float32)


def _back(a):
    """Inverse of _cols for outputs: a view with Julia's axis order."""
    if _is_torch(a):
        return a.permute(*reversed(range(a.dim())))
    return a.transpose(tuple(reversed(range(a.ndim))))


def _layer_descs(model: Chain):
    if model.time_dependent or model.input_activation is not None:
        raise ValueError("mlp_forward/backward take a plain Chain of Dense layers")
    arr = (LayerDesc * len(model.layers))()
    for i, L in enumerate(model.layers):
        arr[i] = LayerDesc(L.in_dims, L.out_dims, _lib.ACT[L.activation])
    return arr


def mlp_forward(model: Chain, ps, x, ctx: Optional[Context] = None):
    """``Chain(Dense...)`` on ``x`` of shape (in, ...): every trailing index is a column (rec_to_gen on
    (2L, B); gen_to_data on the stacked saves (N, T, B), experiments/src/construct.jl:232-246)."""
    ctx = ctx or default_context(0)
    xb = _cols(x)
    N = int(np.prod(xb.shape[:-1]))
    psb = ps.detach().to(torch.float32).contiguous() if _is_torch(ps) else np.ascontiguousarray(np.asarray(ps, np.float32))
    y = _out_like(tuple(xb.shape[:-1]) + (model.layers[-1].out_dims,), xb)
    check(lib().lrnde_mlp_forward(ctx._h, _layer_descs(model), len(model.layers), _ptr(psb), _ptr(xb), N,
                                  0 if _is_torch(x) else 1, _ptr(y)))
    return _back(y)


def mlp_backward(model: Chain, ps, x, d_y, ctx: Optional[Context] = None):
    """(d_x, d_ps) of mlp_forward (the forward is recomputed from ``x``)."""
    ctx = ctx or default_context(0)
    xb, dyb = _cols(x), _cols(d_y)
    N = int(np.prod(xb.shape[:-1]))
    psb = ps.detach().to(torch.float32).contiguous() if _is_torch(ps) else np.ascontiguousarray(np.asarray(ps, np.float32))
    d_x = _out_like(xb.shape, xb)
    d_ps = _out_like((nparams(model),), xb)
    check(lib().lrnde_mlp_backward(ctx._h, _layer_descs(model), len(model.layers), _ptr(psb), _ptr(xb), _ptr(dyb), N,
                                   0 if _is_torch(x) else 1, _ptr(d_x), _ptr(d_ps)))
    return _back(d_x), d_ps


class ReparameterizeLayer:
    """``ReparameterizeLayer()`` (src/layers/common.jl:48-77): x = vcat(mu0, logsigma2) -> mu0 + exp(logsigma2/2) .* eps;
    the state keeps mu0 / logsigma2 for the KL term.  eps comes from Philox(seed) instead of randn_like(rng, ...)."""

    def __init__(self, seed: int = 0, ctx: Optional[Context] = None):
        self.seed, self._ctx = int(seed), ctx

    def initialstates(self, rng: np.random.Generator):
        rng.standard_normal(1)
        return dict(rng=rng, training=True, mu0=None, logsigma2=None)

    def __call__(self, x, ps=None, st=None):
        ctx = self._ctx or default_context(0)
        st = dict(st or {"training": True})
        xb = _cols(x)
        B, L2 = xb.shape
        L = L2 // 2
        y = _out_like((B, L), xb)
        check(lib().lrnde_reparameterize(ctx._h, _ptr(xb), L, B, self.seed, 1 if st.get("training", True) else 0,
                                         0 if _is_torch(x) else 1, _ptr(y), None, None, None, None))
        if st.get("training", True):
            st.update(mu0=x[:L], logsigma2=x[L:])
        else:
            st.update(mu0=x[:L], logsigma2=x[:L])                 # common.jl:73-77
        return _back(y), st

    def backward(self, x, d_y, d_mu=None, d_ls=None, training: bool = True):
        ctx = self._ctx or default_context(0)
        xb = _cols(x)
        B, L2 = xb.shape
        d_x = _out_like(xb.shape, xb)
        f = lambda a: None if a is None else _cols(a)
        bufs = [f(d_y), f(d_mu), f(d_ls)]
        check(lib().lrnde_reparameterize(ctx._h, _ptr(xb), L2 // 2, B, self.seed, 1 if training else 0,
                                         0 if _is_torch(x) else 1, None, _ptr(bufs[0]), _ptr(bufs[1]), _ptr(bufs[2]),
                                         _ptr(d_x)))
        return _back(d_x)


def latent_loss(pred, data, mask, mu, logsigma2, w_kl: float, ctx: Optional[Context] = None):
    """The latent-ODE objective without the regulariser term (experiments/src/construct.jl:36-70):
    returns (loss, neg_log_likelihood, kl_div, d_pred, d_mu, d_logsigma2)."""
    ctx = ctx or default_context(0)
    pb, db, mb, mub, lsb = _cols(pred), _cols(data), _cols(mask), _cols(mu), _cols(logsigma2)
    B, T, F = pb.shape
    L = mub.shape[1]
    out3 = (C.c_float * 3)()
    d_pred, d_mu, d_ls = _out_like(pb.shape, pb), _out_like(mub.shape, pb), _out_like(mub.shape, pb)
    check(lib().lrnde_latent_loss(ctx._h, _ptr(pb), _ptr(db), _ptr(mb), _ptr(mub), _ptr(lsb), F, T, L, B, float(w_kl),
                                  0 if _is_torch(pred) else 1, out3, _ptr(d_pred), _ptr(d_mu), _ptr(d_ls)))
    return float(out3[0]), float(out3[1]), float(out3[2]), _back(d_pred), _back(d_mu), _back(d_ls)


# ------------------------------------------------------------------ torch autograd bridge
if torch is not None:

    class _NeuralODEFn(torch.autograd.Function):
        """Stands in for the ChainRules rrule a Julia wrapper would define: differentiable
        outputs are the stacked saved states and ``reg_val``."""

        @staticmethod
        def forward(ctx, x, ps, layer, st, box):
            sol, st2 = layer(x, ps, st, keep_tape=True)
            box["sol"], box["st"] = sol, st2
            ctx.layer, ctx.sol = layer, sol
            u = torch.stack(sol.u, dim=0)
            reg = torch.tensor(float(st2["reg_val"]), device=x.device, dtype=torch.float32)
            return u, reg

        @staticmethod
        def backward(ctx, d_u, d_reg):
            sol = ctx.sol
            d_us = [d_u[i] for i in range(d_u.shape[0])] if d_u is not None else [None] * len(sol.u)
            d_x, d_ps = ctx.layer.backward(sol, d_us, float(d_reg) if d_reg is not None else 0.0)
            sol.free()
            return d_x, d_ps, None, None, None

    def neural_ode_apply(layer: NeuralODE, x, ps, st):
        """Differentiable call: returns (u_stack [nsave, D, B], reg_val, st')."""
        box = {}
        u, reg = _NeuralODEFn.apply(x, ps, layer, st, box)
        return u, reg, box["st"]


# ------------------------------------------------------------------ layers either side of the cifar10 NeuralODE
def _whcn(a):
    """(W, H, C, B) array in the reference's column-major order == C-contiguous (B, C, H, W) memory."""
    if _is_torch(a):
        return a.detach().to(torch.float32).permute(3, 2, 1, 0).contiguous()
    return np.ascontiguousarray(np.asarray(a, np.float32).transpose(3, 2, 1, 0))


def _from_bchw(a):
    return a.permute(3, 2, 1, 0) if _is_torch(a) else a.transpose(3, 2, 1, 0)


def conv2d_forward(x, ps, out_ch: int, activation: str = "identity", use_bias: bool = True, ctx: Optional[Context] = None):
    """Lux ``Conv((3,3), in => out, activation; pad=(1,1))`` on a WHCN array ``(W, H, in, B)``; ``ps`` = weight
    ``[3,3,in,out]`` (column-major) then bias (experiments/src/construct.jl:220,224)."""
    ctx = ctx or default_context(0)
    xb = _whcn(x)
    B, Cin, H, W = xb.shape
    host = not _is_torch(x)
    psb = np.ascontiguousarray(np.asarray(ps, np.float32)) if host else ps.detach().to(torch.float32).contiguous()
    y = np.empty((B, out_ch, H, W), np.float32) if host else torch.empty((B, out_ch, H, W), dtype=torch.float32, device=xb.device)
    check(lib().lrnde_conv2d_forward(ctx._h, Cin, out_ch, 1 if use_bias else 0, _lib.ACT[activation], W, H, _ptr(psb),
                                     _ptr(xb), B, 1 if host else 0, _ptr(y)))
    return _from_bchw(y)


def conv2d_backward(x, ps, d_y, activation: str = "identity", use_bias: bool = True, need_dx: bool = True,
                    ctx: Optional[Context] = None):
    """Pullback of :func:`conv2d_forward`: ``(d_x or None, d_ps)``."""
    ctx = ctx or default_context(0)
    xb, dyb = _whcn(x), _whcn(d_y)
    B, Cin, H, W = xb.shape
    Cout = dyb.shape[1]
    host = not _is_torch(x)
    psb = np.ascontiguousarray(np.asarray(ps, np.float32)) if host else ps.detach().to(torch.float32).contiguous()
    n_ps = 9 * Cin * Cout + (Cout if use_bias else 0)
    if host:
        d_x = np.empty(xb.shape, np.float32) if need_dx else None
        d_ps = np.empty(n_ps, np.float32)
    else:
        d_x = torch.empty_like(xb) if need_dx else None
        d_ps = torch.empty(n_ps, dtype=torch.float32, device=xb.device)
    check(lib().lrnde_conv2d_backward(ctx._h, Cin, Cout, 1 if use_bias else 0, _lib.ACT[activation], W, H, _ptr(psb),
                                      _ptr(xb), _ptr(dyb), B, 1 if host else 0, _ptr(d_x), _ptr(d_ps)))
    return (_from_bchw(d_x) if need_dx else None), d_ps


def batchnorm_forward(x, ps, state: Optional[dict] = None, activation: str = "identity", ctx: Optional[Context] = None):
    """Lux ``BatchNorm(C, activation)`` on a WHCN array; ``ps`` = scale then bias; ``state`` = ``{"running": [mean; var],
    "training": bool}`` (``Lux.initialstates``: zeros / ones).  Returns ``(y, state')``."""
    ctx = ctx or default_context(0)
    xb = _whcn(x)
    B, Cc, H, W = xb.shape
    host = not _is_torch(x)
    psb = np.ascontiguousarray(np.asarray(ps, np.float32)) if host else ps.detach().to(torch.float32).contiguous()
    y = np.empty(xb.shape, np.float32) if host else torch.empty_like(xb)
    run, test = None, 0
    if state is not None and state.get("running") is not None:
        r = state["running"]
        run = (np.ascontiguousarray(np.asarray(r, np.float32)).copy() if host
               else (r.detach().to(device=xb.device, dtype=torch.float32).clone() if _is_torch(r)
                     else torch.from_numpy(np.asarray(r, np.float32).copy()).to(xb.device)))
        test = 0 if state.get("training", True) else 1
    check(lib().lrnde_batchnorm_forward(ctx._h, Cc, H * W, _lib.ACT[activation], _ptr(psb), _ptr(xb), B, _ptr(run), test,
                                        1 if host else 0, _ptr(y)))
    st2 = state if run is None else dict(state, running=run)
    return _from_bchw(y), st2


def batchnorm_backward(x, ps, d_y, state: Optional[dict] = None, activation: str = "identity", ctx: Optional[Context] = None):
    """Pullback of :func:`batchnorm_forward` (training mode: through the batch statistics): ``(d_x, d_ps)``."""
    ctx = ctx or default_context(0)
    xb, dyb = _whcn(x), _whcn(d_y)
    B, Cc, H, W = xb.shape
    host = not _is_torch(x)
    psb = np.ascontiguousarray(np.asarray(ps, np.float32)) if host else ps.detach().to(torch.float32).contiguous()
    d_x = np.empty(xb.shape, np.float32) if host else torch.empty_like(xb)
    d_ps = np.empty(2 * Cc, np.float32) if host else torch.empty(2 * Cc, dtype=torch.float32, device=xb.device)
    run, test = None, 0
    if state is not None and state.get("running") is not None and not state.get("training", True):
        r = state["running"]
        run = (np.ascontiguousarray(np.asarray(r, np.float32)) if host
               else (r.detach().to(device=xb.device, dtype=torch.float32).contiguous() if _is_torch(r)
                     else torch.from_numpy(np.asarray(r, np.float32).copy()).to(xb.device)))
        test = 1
    check(lib().lrnde_batchnorm_backward(ctx._h, Cc, H * W, _lib.ACT[activation], _ptr(psb), _ptr(xb), _ptr(dyb), B,
                                         _ptr(run), test, 1 if host else 0, _ptr(d_x), _ptr(d_ps)))
    return _from_bchw(d_x), d_ps


class AugmenterLayer:
    """``AugmenterLayer(Conv((3,3), in => extra; pad=(1,1)), 3)``: ``cat(x, conv(x); dims=3)``
    (src/layers/common.jl:80-92, experiments/src/construct.jl:220)."""

    def __init__(self, in_ch: int, extra_ch: int):
        self.in_ch, self.extra_ch = in_ch, extra_ch

    def nparams(self):
        return 9 * self.in_ch * self.extra_ch + self.extra_ch

    def __call__(self, x, ps, ctx=None):
        y = conv2d_forward(x, ps, self.extra_ch, "identity", True, ctx)
        return torch.cat([x, y], dim=2) if _is_torch(x) else np.concatenate([np.asarray(x, np.float32), y], axis=2)

    def backward(self, x, ps, d_out, ctx=None):
        """(d_x, d_ps): the cotangent of the first ``in_ch`` channels passes through, the rest goes through the conv."""
        d_pass, d_conv = d_out[:, :, :self.in_ch], d_out[:, :, self.in_ch:]
        d_x, d_ps = conv2d_backward(x, ps, d_conv, "identity", True, True, ctx)
        return d_x + d_pass, d_ps
