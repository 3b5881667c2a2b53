import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
import oracle as orc
pkg = entry.load_package()
np.set_printoptions(precision=5, linewidth=200, suppress=True)
def run(layers, td, input_act, B):
    om = orc.MLP([orc.Dense(*l) for l in layers], time_dependent=td, input_act=input_act)
    rng = np.random.default_rng(0)
    ps = (orc.glorot_uniform_params(om, rng) * 3 + 0.05 * rng.standard_normal(om.nparams)).astype(np.float32)
    x = rng.standard_normal((layers[0][0], B)).astype(np.float32)
    c = rng.standard_normal((layers[0][0], B)).astype(np.float32)
    ch = pkg.Chain(*[pkg.Dense(*l) for l in layers], input_activation=input_act); ch = pkg.TDChain(ch) if td else ch
    out = {}
    for prec in ("fp32", "tf32x3"):
        node = pkg.NeuralODE(ch, regularize="none", abstol=1e-3, reltol=1e-3, precision=prec)
        sol, st2 = node(x, ps, node.initialstates(np.random.default_rng(0)), keep_tape=True)
        d_x, d_ps = node.backward(sol, [c], 0.0)
        out[prec] = (np.asarray(d_x, np.float64), np.asarray(d_ps, np.float64), sol.bwd_stats.nf_bwd)
    a, b = out["fp32"], out["tf32x3"]
    rel = lambda u, v: np.abs(u - v).max() / (np.abs(v).max() + 1e-30)
    e = np.abs(b[1] - a[1]); 
    print(layers, td, input_act, B, "nf_bwd", a[2], b[2], "d_x rel", f"{rel(b[0], a[0]):.2e}", "d_ps rel", f"{rel(b[1], a[1]):.2e}", "worst idx", int(e.argmax()), flush=True)
run([(16, 12, "tanh"), (12, 16, "identity")], True, None, 8)
run([(40, 200, "tanh"), (200, 40, "identity")], True, None, 100)
run([(20, 40, "tanh"), (40, 20, "tanh"), (20, 40, "tanh"), (40, 20, "tanh")], False, "tanh", 64)
run([(784, 100, "tanh"), (100, 784, "identity")], True, None, 256)
