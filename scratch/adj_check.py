"""Latent-space adjoint (csrc/lrnde_adjoint.cu) against the per-layer adjoint engine on the SAME forward tape.
    python scratch/adj_check.py [B ...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
chain = pkg.TDChain(pkg.Chain(pkg.Dense(784, 100, "tanh"), pkg.Dense(100, 784)))


def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


for B in [int(a) for a in sys.argv[1:]] or [256, 8192]:
    for tol in (1e-3, 1.4e-8):
        node = pkg.NeuralODE(chain, ctx=ctx, abstol=tol, reltol=tol, precision="tf32x3", regularize="unbiased")
        rng = np.random.default_rng(0)
        ps = torch.from_numpy(node.initialparameters(rng) * 3.0).to(dev)
        x = torch.rand((B, 784), device=dev).t()
        st = node.initialstates(np.random.default_rng(1))
        sol, st2 = node(x, ps, st)
        cots = [(torch.randn((B, 784), device=dev) / B).t() for _ in sol.u]
        out = {}
        for mode in ("latent", "layers"):
            if mode == "layers":
                os.environ["LRNDE_NO_LATENT_ADJ"] = "1"
            else:
                os.environ.pop("LRNDE_NO_LATENT_ADJ", None)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d_x, d_ps = node.backward(sol, cots, 0.0)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            s = sol.bwd_stats
            out[mode] = (d_x.clone(), d_ps.clone(), s.naccept_bwd, s.nreject_bwd, ms, s.reserved[3], s.gpu_launches)
            t, dtl, eest, acc = sol.step_log(1)
            print("      dt  ", " ".join(f"{v:.6e}" for v in dtl[:6]))
            print("      EEst", " ".join(f"{v:.6e}" for v in eest[:6]))
            print(f"B={B} tol={tol:g} {mode:7s}: bwd steps {s.naccept_bwd}+{s.nreject_bwd} nf {s.nf_bwd} ret {s.retcode_bwd} "
                  f"wall {ms:.2f} ms adjoint {s.reserved[3]} us launches {s.gpu_launches} |d_ps| {float(d_ps.abs().max()):.4e}")
        a, b = out["latent"], out["layers"]
        print(f"   latent vs layers: d_x rel {rel(a[0], b[0]):.3e}  d_ps rel {rel(a[1], b[1]):.3e}  "
              f"speedup (adjoint phase) {b[5] / max(a[5], 1):.2f}x")
os.environ.pop("LRNDE_NO_LATENT_ADJ", None)
