"""In-kernel timeline of CTA 0 of the LAST tensor-core convolution of one f evaluation (LRNDE_CT_DBG=32)."""
import os, sys, ctypes as C, numpy as np, torch
os.environ["LRNDE_CT_DBG"] = os.environ.get("LRNDE_CT_DBG", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
layers = [pkg.Conv(8, 64, True, "gelu")] if (len(sys.argv) > 2 and sys.argv[2] == "k8") else \
         [pkg.Conv(8, 64, True, "gelu"), pkg.Conv(64, 64, True, "gelu")]
# dynamics must map the state onto itself: append a plain conv back to 8 channels and trace through an env-chosen layer
layers.append(pkg.Conv(64, 8))
chain = pkg.TDConvChain(pkg.ConvChain(*layers, width=32, height=32))
layer = pkg.NeuralODE(chain)
ps = torch.from_numpy(layer.initialparameters(np.random.default_rng(0))).cuda()
u = torch.randn((8192, B), device="cuda")
if len(sys.argv) > 3 and sys.argv[3] == "vjp":   # the last matching convolution of a right-hand side (data gradient, pullback epilogue)
    un, pn = u.cpu().numpy(), ps.cpu().numpy()
    for _ in range(3): layer.dynamics_vjp(un, pn, 0.5, un)
else:
    for _ in range(3): layer.dynamics(u, ps, 0.5)
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
lib.lrnde_debug_trace_convtc(buf, 128)
t = np.array(buf[:]).reshape(8, 16)
t0 = t[0, 0]
names = ["setup done", "producer issues group", "mma sees stage", "mma committed", "epi waits", "epi starts", "epi released tmem"]
for r, nm in enumerate(names):
    print(f"{nm:24s}", " ".join(f"{(x - t0) / 1.9:9.0f}" if x else "        -" for x in t[r, :5]), " (ns at 1.9 GHz)")
