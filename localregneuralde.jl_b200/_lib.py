"""ctypes binding of libLRNDE.so (include/lrnde.h).  No fallback: a missing library or a
missing CUDA device is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libLRNDE.so")

ACT = {"identity": 0, None: -1, "tanh": 1, "gelu": 2, "sigmoid": 3, "relu": 4}
REG = {"none": 0, "unbiased": 1, "biased": 2}
REGTYPE = {"error_estimate": 0, "stiffness_estimate": 1}
PREC = {"auto": 0, "fp32": 1, "tf32x3": 2, "tf32": 3, "smem": 4}
POW = {"fastpow_2023": 0, "exact": 1}
RETCODES = {0: "Success", 1: "MaxIters", 2: "DtLessThanMin", 3: "Unstable", 4: "TapeFull", 5: "PeerTimeout"}


class LayerDesc(C.Structure):
    _fields_ = [("in_dims", C.c_int32), ("out_dims", C.c_int32), ("act", C.c_int32)]


class ConvLayerDesc(C.Structure):
    _fields_ = [("in_ch", C.c_int32), ("out_ch", C.c_int32), ("batchnorm", C.c_int32), ("act", C.c_int32)]


class Opts(C.Structure):
    _fields_ = [
        ("t0", C.c_float), ("t2", C.c_float), ("abstol", C.c_float), ("reltol", C.c_float),
        ("maxiters", C.c_int32), ("reg_mode", C.c_int32), ("reg_type", C.c_int32),
        ("t1", C.c_float), ("u01", C.c_float), ("saveat", C.POINTER(C.c_float)),
        ("nsave", C.c_int32), ("save_start", C.c_int32), ("precision", C.c_int32),
        ("pow_mode", C.c_int32), ("host_buffers", C.c_int32), ("keep_tape", C.c_int32),
        ("loop_mode", C.c_int32), ("last_only", C.c_int32), ("model_state", C.c_void_p),
        ("model_testmode", C.c_int32), ("no_dx", C.c_int32), ("reserved", C.c_int32 * 2),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("nfe", C.c_int32), ("naccept", C.c_int32), ("nreject", C.c_int32),
        ("retcode", C.c_int32), ("reg_val", C.c_float), ("t1_used", C.c_float),
        ("dt_reg", C.c_float), ("nsave_out", C.c_int32), ("nf_bwd", C.c_int32),
        ("naccept_bwd", C.c_int32), ("nreject_bwd", C.c_int32), ("retcode_bwd", C.c_int32),
        ("gpu_launches", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class SdeOpts(C.Structure):
    _fields_ = [
        ("t0", C.c_float), ("t2", C.c_float), ("abstol", C.c_float), ("reltol", C.c_float),
        ("maxiters", C.c_int32), ("reg_mode", C.c_int32), ("t1", C.c_float), ("u01", C.c_float),
        ("saveat", C.POINTER(C.c_float)), ("nsave", C.c_int32), ("save_start", C.c_int32),
        ("seed", C.c_uint64), ("pow_mode", C.c_int32), ("host_buffers", C.c_int32),
        ("keep_tape", C.c_int32), ("qmax", C.c_float), ("gamma", C.c_float), ("qmin", C.c_float),
        ("delta", C.c_float), ("reserved", C.c_int32 * 8),
    ]


class SdeStats(C.Structure):
    _fields_ = [
        ("nfe_drift", C.c_int32), ("nfe_diffusion", C.c_int32), ("naccept", C.c_int32),
        ("nreject", C.c_int32), ("retcode", C.c_int32), ("reg_val", C.c_float),
        ("t1_used", C.c_float), ("dt_reg", C.c_float), ("nsave_out", C.c_int32),
        ("ndraws", C.c_int32), ("gpu_launches", C.c_int32), ("reserved", C.c_int32 * 8),
    ]


class LrndeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libLRNDE error {code}: {msg}")
        self.code = code


_lib = None

# every symbol include/lrnde.h declares (tests check the library exports all of them)
SYMBOLS = [
    "lrnde_last_error", "lrnde_version", "lrnde_ctx_create", "lrnde_ctx_destroy",
    "lrnde_ctx_sync", "lrnde_ctx_set_tape_budget", "lrnde_ctx_mailbox", "lrnde_ctx_set_dist",
    "lrnde_model_create", "lrnde_model_destroy", "lrnde_model_nparams",
    "lrnde_model_state_dims", "lrnde_dynamics_eval", "lrnde_ode_forward", "lrnde_ode_backward",
    "lrnde_tape_free", "lrnde_step_log", "lrnde_sosri_step", "lrnde_ipc_export",
    "lrnde_ipc_open", "lrnde_head_ce", "lrnde_adam_step", "lrnde_profile_feval",
    "lrnde_sde_forward", "lrnde_sde_backward", "lrnde_sde_tape_free", "lrnde_sde_states",
    "lrnde_sde_step_log", "lrnde_sde_aux_step", "lrnde_gru_nparams", "lrnde_gru_forward",
    "lrnde_gru_backward", "lrnde_gru_tape_free", "lrnde_mlp_forward", "lrnde_mlp_backward",
    "lrnde_reparameterize", "lrnde_latent_loss", "lrnde_conv_model_create", "lrnde_dynamics_vjp",
    "lrnde_conv2d_forward", "lrnde_conv2d_backward", "lrnde_batchnorm_forward", "lrnde_batchnorm_backward",
    "lrnde_opt_step", "lrnde_allreduce_sum", "lrnde_ode_saved_states", "lrnde_classifier_grad", "lrnde_prefetch_inputs",
    "lrnde_profile_adjoint_last", "lrnde_profile_step",
]

OPT = {"descent": 0, "momentum": 1, "nesterov": 2, "adam": 3, "adamax": 4}


def lib():
    """Loads libLRNDE.so (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LrndeError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                             "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.lrnde_last_error.restype = C.c_char_p
    L.lrnde_version.restype = i32
    L.lrnde_ctx_create.argtypes = [C.POINTER(vp), i32, vp]
    L.lrnde_ctx_destroy.argtypes = [vp]
    L.lrnde_ctx_sync.argtypes = [vp]
    L.lrnde_ctx_set_tape_budget.argtypes = [vp, C.c_uint64]
    L.lrnde_ctx_mailbox.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.lrnde_ctx_set_dist.argtypes = [vp, i32, i32, C.POINTER(vp), i64]
    L.lrnde_model_create.argtypes = [vp, C.POINTER(LayerDesc), i32, i32, i32, C.POINTER(vp)]
    L.lrnde_model_destroy.argtypes = [vp]
    L.lrnde_conv_model_create.argtypes = [vp, C.POINTER(ConvLayerDesc), i32, i32, i32, i32, C.POINTER(vp)]
    L.lrnde_dynamics_vjp.argtypes = [vp, vp, C.POINTER(Opts), vp, vp, f32, vp, i64, vp, vp]
    L.lrnde_conv2d_forward.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, i64, i32, vp]
    L.lrnde_conv2d_backward.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i64, i32, vp, vp]
    L.lrnde_batchnorm_forward.argtypes = [vp, i32, i64, i32, vp, vp, i64, vp, i32, i32, vp]
    L.lrnde_batchnorm_backward.argtypes = [vp, i32, i64, i32, vp, vp, vp, i64, vp, i32, i32, vp, vp]
    L.lrnde_model_nparams.argtypes = [vp]
    L.lrnde_model_nparams.restype = i64
    L.lrnde_model_state_dims.argtypes = [vp]
    L.lrnde_model_state_dims.restype = i64
    L.lrnde_dynamics_eval.argtypes = [vp, vp, C.POINTER(Opts), vp, vp, f32, i64, vp]
    L.lrnde_ode_forward.argtypes = [vp, vp, C.POINTER(Opts), vp, vp, i64, vp, i64, vp,
                                    C.POINTER(Stats), C.POINTER(vp)]
    L.lrnde_ode_backward.argtypes = [vp, vp, vp, vp, f32, vp, vp, C.POINTER(Stats)]
    L.lrnde_tape_free.argtypes = [vp]
    L.lrnde_step_log.argtypes = [vp, i32, vp, vp, vp, vp, i32, C.POINTER(i32)]
    L.lrnde_sosri_step.argtypes = [vp, vp, vp, C.POINTER(Opts), vp, vp, vp, vp, vp, f32, f32,
                                   f32, i64, vp, C.POINTER(f32)]
    L.lrnde_ipc_export.argtypes = [vp, vp, vp]
    L.lrnde_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.lrnde_head_ce.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, vp, vp, vp]
    L.lrnde_adam_step.argtypes = [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, i32]
    L.lrnde_profile_feval.argtypes = [vp, vp, C.POINTER(Opts), vp, vp, i64, i32, vp,
                                      C.POINTER(f32), C.POINTER(i32)]
    L.lrnde_sde_forward.argtypes = [vp, vp, vp, C.POINTER(SdeOpts), vp, vp, vp, i64, vp,
                                    C.POINTER(SdeStats), C.POINTER(vp)]
    L.lrnde_sde_backward.argtypes = [vp, vp, vp, f32, vp, vp, vp]
    L.lrnde_sde_tape_free.argtypes = [vp]
    L.lrnde_sde_states.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.lrnde_sde_step_log.argtypes = [vp, vp, vp, vp, vp, i32, C.POINTER(i32)]
    L.lrnde_sde_aux_step.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, f32, f32, f32, f32, f32, i64, i32, vp,
                                     C.POINTER(f32)]
    L.lrnde_gru_nparams.argtypes = [i32, i32, i32]
    L.lrnde_gru_nparams.restype = i64
    L.lrnde_gru_forward.argtypes = [vp, i32, i32, i32, vp, vp, i32, i64, i32, i32, vp, C.POINTER(vp)]
    L.lrnde_gru_backward.argtypes = [vp, vp, vp, vp]
    L.lrnde_gru_tape_free.argtypes = [vp]
    L.lrnde_mlp_forward.argtypes = [vp, C.POINTER(LayerDesc), i32, vp, vp, i64, i32, vp]
    L.lrnde_mlp_backward.argtypes = [vp, C.POINTER(LayerDesc), i32, vp, vp, vp, i64, i32, vp, vp]
    L.lrnde_reparameterize.argtypes = [vp, vp, i32, i64, C.c_uint64, i32, i32, vp, vp, vp, vp, vp]
    L.lrnde_latent_loss.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i64, f32, i32, vp, vp, vp, vp]
    L.lrnde_opt_step.argtypes = [vp, i32, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32]
    L.lrnde_allreduce_sum.argtypes = [vp, vp, i64]
    L.lrnde_ode_saved_states.argtypes = [vp, vp, vp, vp, i64, vp]
    L.lrnde_prefetch_inputs.argtypes = [vp, vp, vp, i64, i32]
    L.lrnde_classifier_grad.argtypes = [vp, vp, C.POINTER(Opts), vp, vp, vp, vp, i64, i32, f32, f32, vp, vp, vp,
                                        C.POINTER(Stats)]
    L.lrnde_profile_adjoint_last.argtypes = [vp]
    L.lrnde_profile_step.argtypes = [vp, vp, C.POINTER(Opts), vp, vp, i64, i32, vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("lrnde_last_error", "lrnde_model_nparams", "lrnde_model_state_dims", "lrnde_gru_nparams"):
            fn.restype = i32
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise LrndeError(rc, lib().lrnde_last_error().decode("utf-8", "replace"))
