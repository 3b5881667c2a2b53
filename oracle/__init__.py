"""CPU oracle for the LocalRegNeuralDE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product path
(``localregneuralde_b200``) never does; it fails loudly when libLRNDE.so is
missing instead of falling back to anything here.

PARITY UNPINNED: the reference (Julia) cannot run in this image and its own
tests hold no golden vectors (test/runtests.jl:21-29,118-131 assert only
finiteness / non-zeroness).  See oracle/lrnde_oracle.py's header.
"""
from .lrnde_oracle import *  # noqa: F401,F403
from .lrnde_sde_oracle import *  # noqa: F401,F403,E402
from .lrnde_latent_oracle import *  # noqa: F401,F403,E402
from .lrnde_conv_oracle import *  # noqa: F401,F403,E402
