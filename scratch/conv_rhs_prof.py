"""One f evaluation and one adjoint right-hand side (dynamics_vjp) of the cifar10 node core at batch B, device-resident
inputs.  Under ncu:  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv \
    python scratch/conv_rhs_prof.py 256 1      (the LAST call of each kind is the warm one)
Without ncu: CUDA-event time of `reps` back-to-back calls (includes the per-call setup of the C ABI)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
chain = pkg.TDConvChain(pkg.ConvChain(pkg.Conv(8, 64, True, "gelu"), pkg.Conv(64, 64, True, "gelu"), pkg.Conv(64, 8), width=32, height=32))
layer = pkg.NeuralODE(chain)
rng = np.random.default_rng(0)
ps = torch.from_numpy(layer.initialparameters(rng)).to(dev)
u = torch.randn((8192, B), device=dev).t().contiguous().t() if False else torch.randn((B, 8192), device=dev).t()
u = torch.randn((8192, B), device=dev)
lam = torch.randn((8192, B), device=dev)
def timeit(fn):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps
print("f    us", timeit(lambda: layer.dynamics(u, ps, 0.5)))
un, pn, ln = u.cpu().numpy(), ps.cpu().numpy(), lam.cpu().numpy()
print("vjp  us (host buffers)", timeit(lambda: layer.dynamics_vjp(un, pn, 0.5, ln)))
