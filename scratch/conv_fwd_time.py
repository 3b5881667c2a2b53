"""Forward pass of the cifar10 node core (batch 256): CUDA-event time per call, with and without the regulariser."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
B = 256
chain = pkg.TDConvChain(pkg.ConvChain(pkg.Conv(8, 64, True, "gelu"), pkg.Conv(64, 64, True, "gelu"), pkg.Conv(64, 8), width=32, height=32))
x = torch.from_numpy(np.random.default_rng(1).standard_normal((8192, B)).astype(np.float32)).cuda()
for mode in ("none", "unbiased"):
    for lm in (0, 2):
        layer = pkg.NeuralODE(chain, return_last_only=True, loop_mode=lm, regularize=mode, abstol=1e-4, reltol=1e-4, save_start=False, maxiters=10000)
        ps = torch.from_numpy(layer.initialparameters(np.random.default_rng(0))).cuda()
        st = layer.initialstates(np.random.default_rng(7))
        for _ in range(3):
            sol, st2 = layer(x, ps, st); sol.free()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            sol, st2 = layer(x, ps, st); nfe = st2["nfe"]; s = sol.stats; sol.free()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"regularize={mode:9s} loop_mode={lm}: {ms:7.2f} ms per forward pass, nfe {nfe}, accepted {s.naccept} rejected {s.nreject}, launches {s.gpu_launches} -> {1e3 * ms / nfe:.0f} us per evaluation")
