"""Hidden-space reverse pass of the regulariser step (lrnde_regrev.cuh) against the layer-by-layer one
(LRNDE_NO_HIDDEN_REG=1), same forward solve: d reg / d ps only (zero cotangent on the states)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
for (D, H, B, td, tol, scale) in ((784, 100, 8192, True, 1.4e-8, 1.0), (784, 100, 128, True, 1e-5, 3.0), (64, 40, 77, False, 1e-4, 2.0)):
    inner = pkg.Chain(pkg.Dense(D, H, "tanh"), pkg.Dense(H, D))
    chain = pkg.TDChain(inner) if td else inner
    node = pkg.NeuralODE(chain, ctx=ctx, abstol=tol, reltol=tol, precision="tf32x3", regularize="unbiased", save_start=False)
    ps = torch.from_numpy(node.initialparameters(np.random.default_rng(0)) * scale).to(dev)
    x = torch.rand((B, D), device=dev).t()
    res = {}
    for mode in ("layer", "hidden", "hidden"):
        if mode == "layer": os.environ["LRNDE_NO_HIDDEN_REG"] = "1"
        else: os.environ.pop("LRNDE_NO_HIDDEN_REG", None)
        sol, st2 = node(x, ps, node.initialstates(np.random.default_rng(1)))
        zero = torch.zeros((B, D), device=dev).t()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d_x, d_ps = node.backward(sol, [None, zero], 1.0)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res[mode] = (d_ps.cpu().numpy().copy(), sol.bwd_stats.reserved[4], float(st2["reg_val"]))
        sol.free()
    a, b = res["layer"][0], res["hidden"][0]
    err = np.abs(a - b).max() / (np.abs(a).max() + 1e-30)
    P1 = H * (D + (2 if td else 1))
    e1 = np.abs(a[:P1] - b[:P1]).max() / (np.abs(a[:P1]).max() + 1e-30)
    e2 = np.abs(a[P1:] - b[P1:]).max() / (np.abs(a[P1:]).max() + 1e-30)
    print(f"D={D} H={H} B={B} td={td}: reg {res['layer'][2]:.4e}  |d_ps| {np.abs(a).max():.3e}  rel err {err:.2e} (layer 1 {e1:.2e}, layer 2 {e2:.2e})  "
          f"reg_pullback us: layer {res['layer'][1]} hidden {res['hidden'][1]}", flush=True)
