"""In-kernel timeline of CTA 0 of the last weight-gradient kernel of one right-hand side (LRNDE_WG_DBG=32, +64: NB = 64 only)."""
import os, sys, ctypes as C, numpy as np, torch
os.environ["LRNDE_WG_DBG"] = os.environ.get("LRNDE_WG_DBG", "96")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.lib()
B = 256
chain = pkg.TDConvChain(pkg.ConvChain(pkg.Conv(8, 64, True, "gelu"), pkg.Conv(64, 64, True, "gelu"), pkg.Conv(64, 8), width=32, height=32))
layer = pkg.NeuralODE(chain)
ps = layer.initialparameters(np.random.default_rng(0))
u = np.random.default_rng(1).standard_normal((8192, B)).astype(np.float32)
for _ in range(3): layer.dynamics_vjp(u, ps, 0.5, u)
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
lib.lrnde_debug_trace_convtc(buf, 128)
t = np.array(buf[:]).reshape(8, 16); t0 = t[0, 0]
for r, nm in enumerate(["setup done", "TMA of P row issued", "P worker sees the tile", "P worker done", "MMA issue starts", "epilogue: waits / starts / ends"]):
    print(f"{nm:24s}", " ".join(f"{(x - t0) / 1.9:7.0f}" if x else "      -" for x in t[r, :12]), " (ns)")
