import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
import oracle as orc
pkg = entry.load_package()
np.set_printoptions(precision=4, linewidth=220, suppress=True)
layers, td, B = [(16, 12, "tanh"), (12, 16, "identity")], True, 8
om = orc.MLP([orc.Dense(*l) for l in layers], time_dependent=td)
rng = np.random.default_rng(0)
ps = (orc.glorot_uniform_params(om, rng) * 3 + 0.05 * rng.standard_normal(om.nparams)).astype(np.float32)
x = rng.standard_normal((16, B)).astype(np.float32); c = rng.standard_normal((16, B)).astype(np.float32)
ch = pkg.TDChain(pkg.Chain(*[pkg.Dense(*l) for l in layers]))
res = {}
for prec in ("fp32", "tf32x3"):
    node = pkg.NeuralODE(ch, regularize="none", abstol=1e-3, reltol=1e-3, precision=prec, maxiters=1)
    sol, st2 = node(x, ps, node.initialstates(np.random.default_rng(0)), keep_tape=True)
    d_x, d_ps = node.backward(sol, [c], 0.0)
    res[prec] = np.asarray(d_ps, np.float64)
a, b = res["fp32"], res["tf32x3"]
W1 = slice(0, 12 * 17); b1 = slice(12 * 17, 12 * 17 + 12); o2 = 12 * 17 + 12; W2 = slice(o2, o2 + 16 * 13); b2 = slice(o2 + 16 * 13, o2 + 16 * 13 + 16)
print("dW1 fp32\n", a[W1].reshape((12, 17), order="F")[:4, :8], "\ndW1 umma\n", b[W1].reshape((12, 17), order="F")[:4, :8])
print("db1", a[b1][:6], b[b1][:6])
print("dW2 fp32\n", a[W2].reshape((16, 13), order="F")[:4, :8], "\ndW2 umma\n", b[W2].reshape((16, 13), order="F")[:4, :8])
print("db2", a[b2][:6], b[b2][:6])
