import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.lib()
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
L = C.CDLL(pkg.LIB_PATH)
names = ["loader_issue", "mma_full_ok", "mma_issued", "prod_empty_ok", "prod_stored", "prod_fenced", "epi", "start"]
def run(chain, B, label, iters=20):
    node = pkg.NeuralODE(chain, ctx=ctx)
    ps = torch.from_numpy(node.initialparameters(np.random.default_rng(0))).to(dev)
    x = torch.rand((B, 784), device=dev); du = torch.empty_like(x)
    o, _ = node._opts("none", 0.0, 0.0, False, False)
    ms, lp = C.c_float(), C.c_int32()
    pkg._lib.check(lib.lrnde_profile_feval(ctx._h, ctx.model_handle(chain), C.byref(o), ps.data_ptr(), x.data_ptr(), B, iters, du.data_ptr(), C.byref(ms), C.byref(lp)))
    print(f"== {label} B={B} ms/eval={ms.value:.4f} launches={lp.value}")
    if hasattr(L, "lrnde_debug_trace"):
        buf = (C.c_longlong * 1024)()
        L.lrnde_debug_trace(buf, 1024)
        for k, nm in enumerate(("RING layer1", "RESIDENT layer2")):
            a = np.array(buf[:])[k*512:(k+1)*512].reshape(8, 64)
            t0 = a[7, 0]
            print(nm)
            for i, nme in enumerate(names):
                print(f"{nme:14s}", " ".join(f"{int(v - t0):6d}" if v else "     ." for v in a[i][:32]))
l1 = pkg.TDChain(pkg.Chain(pkg.Dense(784, 100, "tanh")))
full = pkg.TDChain(pkg.Chain(pkg.Dense(784, 100, "tanh"), pkg.Dense(100, 784)))
for B in (8192, 2048):
    run(full, B, "full (trace = layer 2 for it<28)")
