"""One forward + one latent-space adjoint with loop_mode=2 (no CUDA graph: every kernel visible to ncu).
    python scratch/adj_prof.py [B] [scale]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
chain = pkg.TDChain(pkg.Chain(pkg.Dense(784, 100, "tanh"), pkg.Dense(100, 784)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
node = pkg.NeuralODE(chain, ctx=ctx, abstol=1.4e-8, reltol=1.4e-8, precision="tf32x3", regularize="unbiased",
                     loop_mode=int(os.environ.get("LOOP_MODE", "2")))
ps = torch.from_numpy(node.initialparameters(np.random.default_rng(0)) * scale).to(dev)
x = torch.rand((B, 784), device=dev).t()
for it in range(int(os.environ.get("ITERS", "2"))):
    sol, st2 = node(x, ps, node.initialstates(np.random.default_rng(1)))
    cots = [(torch.randn((B, 784), device=dev) / B).t() for _ in sol.u]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d_x, d_ps = node.backward(sol, cots, 0.0)
    torch.cuda.synchronize()
    s = sol.bwd_stats
    print(f"B={B} fwd steps {sol.stats.naccept}+{sol.stats.nreject} ({sol.stats.reserved[0]} us)  bwd steps {s.naccept_bwd}+{s.nreject_bwd} "
          f"adjoint {s.reserved[3]} us  wall {(time.perf_counter() - t0) * 1e3:.2f} ms launches {s.gpu_launches}")
