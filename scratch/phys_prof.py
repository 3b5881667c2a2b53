import os, sys, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry, oracle as orc
pkg = entry.load_package()
Nn, H, T, B = 20, 40, 49, 256
rng = np.random.default_rng(8)
dyn = [(Nn, H, "tanh"), (H, Nn, "tanh")] * 4
om = orc.MLP([orc.Dense(*l) for l in dyn], time_dependent=False, input_act="tanh")
ps = (orc.glorot_uniform_params(om, rng) * 2).astype(np.float32)
ts = np.sort(np.concatenate([[0.0], rng.uniform(0.02, 1.0, T - 1)])).astype(np.float32)
lm = int(sys.argv[1]) if len(sys.argv) > 1 else 0
node = pkg.NeuralODE(pkg.Chain(*[pkg.Dense(*l) for l in dyn], input_activation="tanh"), regularize="unbiased", abstol=1e-4,
                     reltol=1e-4, maxiters=10000, saveat=list(ts), loop_mode=lm)
x = rng.standard_normal((Nn, B)).astype(np.float32)
cots = [(rng.standard_normal((Nn, B)) / B).astype(np.float32) for _ in range(T)]
for it in range(3):
    t0 = time.perf_counter()
    sol, st = node(x, ps, node.initialstates(np.random.default_rng(2)))
    t1 = time.perf_counter()
    node.backward(sol, cots, 0.5)
    t2 = time.perf_counter()
    print(f"fwd {1e3*(t1-t0):.2f} ms bwd {1e3*(t2-t1):.2f} ms launches fwd {sol.stats.gpu_launches} bwd {sol.bwd_stats.gpu_launches} steps_bwd {sol.bwd_stats.naccept_bwd}")
    sol.free()
