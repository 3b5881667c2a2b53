import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for B in (128, 8192):
    layer, ps, x = bench.sde_setup(B, False)
    dev = torch.device("cuda", 0)
    xt, pt = torch.from_numpy(x).to(dev), torch.from_numpy(ps).to(dev)
    cot_t = torch.ones((32, B), device=dev) / B; cot_h = np.ones((32, B), np.float32) / B
    st = layer.initialstates(np.random.default_rng(7))
    for resident in (True, False, True, False):
        tf, tb = [], []
        for i in range(8):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            sol, st2 = layer(xt if resident else x, pt if resident else ps, st)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            layer.backward(sol, [None, cot_t if resident else cot_h], 2.5)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            sol.free()
            tf.append(t1 - t0); tb.append(t2 - t1)
        print(f"B={B} resident={resident} fwd ms {[round(1e3*v,2) for v in tf]} bwd ms {[round(1e3*v,2) for v in tb]} attempts {sol.stats.naccept + sol.stats.nreject}")
