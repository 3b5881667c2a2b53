"""B200-native hot path of LocalRegNeuralDE.jl behind the reference's Lux-style layer API.

The directory name contains a dot, so it is loaded through ``__graft_entry__.load_package()``
(importlib) under the module name ``lrnde_b200``.  Everything numerical lives in
``libLRNDE.so`` (csrc/, C ABI in include/lrnde.h); there is no CPU fallback.
"""
from ._lib import LIB_PATH, LrndeError, lib, SYMBOLS  # noqa: F401
from .layers import (AugmenterLayer, Chain, Context, Conv, ConvChain, Dense, DESolution, LatentGRUCell, NeuralDSDE, NeuralODE, Recurrence,  # noqa: F401
                     ReparameterizeLayer, SDESolution, TDChain, TDConvChain, latent_loss, mlp_backward, mlp_forward,
                     default_context, diffeqsol_to_array, diffeqsol_to_timeseries,
                     glorot_uniform, nparams, conv2d_forward, conv2d_backward, batchnorm_forward,
                     batchnorm_backward, initial_model_state)

try:
    from .layers import neural_ode_apply  # noqa: F401
except ImportError:  # torch missing
    pass
from . import trainer  # noqa: F401,E402
