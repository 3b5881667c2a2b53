import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.lib()
D, H, B = 784, 100, 8192
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
chain = pkg.TDChain(pkg.Chain(pkg.Dense(D, H, "tanh"), pkg.Dense(H, D)))
node = pkg.NeuralODE(chain, regularize="none", abstol=1e-3, reltol=1e-3, maxiters=10000, ctx=ctx)
ps = torch.from_numpy(node.initialparameters(np.random.default_rng(0))).to(dev)
x = torch.rand((B, D), device=dev)
st = node.initialstates(np.random.default_rng(7))
sol, st2 = node(x.t(), ps, st, keep_tape=True)
d_x, d_ps = node.backward(sol, [torch.ones((D, B), device=dev) / B], 0.0)
torch.cuda.synchronize()
L = C.CDLL(pkg.LIB_PATH)
buf = (C.c_longlong * 1536)()
L.lrnde_debug_trace(buf, 1536)
a = np.array(buf[:])[1024:1536].reshape(8, 64)
t0 = a[7, 0]
names = ["-", "mma_full_ok", "mma_issued", "prod_empty_ok", "prod_stored", "-", "epi", "start"]
for i, nme in enumerate(names):
    if nme == "-": continue
    print(f"{nme:14s}", " ".join(f"{int(v - t0):6d}" if v else "     ." for v in a[i][:16]))
