import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
chain = pkg.TDChain(pkg.Chain(pkg.Dense(784, 100, "tanh"), pkg.Dense(100, 784)))
node = pkg.NeuralODE(chain, ctx=ctx, precision=(sys.argv[3] if len(sys.argv) > 3 else "auto"))
ps = torch.from_numpy(node.initialparameters(np.random.default_rng(0))).to(dev)
x = torch.rand((B, 784), device=dev); du = torch.empty_like(x)
o, _ = node._opts("none", 0.0, 0.0, False, False)
ms, lp = C.c_float(), C.c_int32()
pkg._lib.check(lib.lrnde_profile_feval(ctx._h, ctx.model_handle(chain), C.byref(o), ps.data_ptr(), x.data_ptr(), B, iters, du.data_ptr(), C.byref(ms), C.byref(lp)))
print("B", B, "ms/feval", ms.value, "TFLOP/s", 2.0*B*(100*785+784*101)/ms.value/1e9)
