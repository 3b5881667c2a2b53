import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.lib()
B = 8192
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
chain = pkg.TDChain(pkg.Chain(pkg.Dense(784, 100, "tanh"), pkg.Dense(100, 784)))
node = pkg.NeuralODE(chain, ctx=ctx)
ps = torch.from_numpy(node.initialparameters(np.random.default_rng(0))).to(dev)
x = torch.rand((B, 784), device=dev); du = torch.empty_like(x)
o, _ = node._opts("none", 0.0, 0.0, False, False)
ms, lp = C.c_float(), C.c_int32()
L = C.CDLL(pkg.LIB_PATH)
for which in (0, 1):
    # run f-eval twice; trace is overwritten by the LAST kernel that ran with these block ids:
    # layer-1 (ring) then layer-2 (resident).  To see layer 1, use a 1-layer chain.
    ch = chain if which == 1 else pkg.TDChain(pkg.Chain(pkg.Dense(784, 100, "tanh"), pkg.Dense(100, 784)))
    pkg._lib.check(lib.lrnde_profile_feval(ctx._h, ctx.model_handle(chain), C.byref(o), ps.data_ptr(), x.data_ptr(), B, 2, du.data_ptr(), C.byref(ms), C.byref(lp)))
buf = (C.c_longlong * 512)()
L.lrnde_debug_trace(buf, 512)
a = np.array(buf[:]).reshape(8, 64)
t0 = a[7, 0]
names = ["loader_issue", "mma_full_ok", "mma_issued", "prod_empty_ok", "prod_stored", "prod_fenced", "epi", "start"]
print("resident kernel (last run) trace, cycles since tmem alloc done")
for i, nme in enumerate(names):
    row = a[i]
    print(f"{nme:14s}", " ".join(f"{int(v - t0):6d}" if v else "     ." for v in row[:30]))
