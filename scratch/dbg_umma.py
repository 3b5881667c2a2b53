import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
import oracle as orc
pkg = entry.load_package()
np.set_printoptions(precision=5, linewidth=220, suppress=True)
def run(layers, td, input_act, B, seed=0):
    om = orc.MLP([orc.Dense(*l) for l in layers], time_dependent=td, input_act=input_act)
    rng = np.random.default_rng(seed)
    ps = orc.glorot_uniform_params(om, rng) + (0.05 * rng.standard_normal(om.nparams)).astype(np.float32)
    x = rng.standard_normal((layers[0][0], B)).astype(np.float32)
    c = pkg.Chain(*[pkg.Dense(*l) for l in layers], input_activation=input_act)
    c = pkg.TDChain(c) if td else c
    want = om.f(x.astype(np.float64), ps.astype(np.float64), 0.37)
    out = {}
    for prec in ("fp32", "tf32x3", "tf32"):
        got = np.asarray(pkg.NeuralODE(c, precision=prec).dynamics(x, ps, 0.37), np.float64)
        err = np.abs(got - want)
        out[prec] = err.max() / np.abs(want).max()
        if prec != "fp32" and out[prec] > 1e-2:
            bad = np.argwhere(err > 1e-2 * np.abs(want).max())
            print("   BAD", prec, "count", len(bad), "of", err.size, "first", bad[:6].tolist(), "rows(feat) bad", sorted(set(bad[:,0].tolist()))[:12], "cols(samp) bad", sorted(set(bad[:,1].tolist()))[:12])
            print("   got", got[:4,:6].ravel(), "\n   want", want[:4,:6].ravel())
    print(layers, td, input_act, B, {k: f"{v:.2e}" for k, v in out.items()}, flush=True)
run([(2, 4, "gelu"), (4, 2, "identity")], True, None, 1)
run([(32, 32, "tanh"), (32, 32, "identity")], False, None, 64)
run([(32, 32, "tanh"), (32, 32, "identity")], True, None, 64)
run([(3, 5, "tanh"), (5, 3, "tanh")], False, "tanh", 7)
run([(64, 128, "tanh"), (128, 64, "identity")], True, None, 128)
run([(200, 130, "tanh"), (130, 200, "identity")], True, None, 100)
run([(784, 100, "tanh"), (100, 784, "identity")], True, None, 128)
run([(784, 100, "tanh"), (100, 784, "identity")], True, None, 77)
run([(20, 40, "tanh"), (40, 20, "tanh"), (20, 40, "sigmoid"), (40, 20, "relu")], True, None, 130)
