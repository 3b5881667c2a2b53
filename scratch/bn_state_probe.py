"""Which variant of the oracle's closure-call sequence does the GPU's returned BatchNorm state match?"""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as entry
import oracle as orc
from oracle.lrnde_conv_oracle import ConvLayer, ConvNet, glorot_uniform_conv_params, initial_conv_state
pkg = entry.load_package()
layers, W, H, B = [(2, 6, True, "gelu"), (6, 20, True, "gelu"), (20, 2, False, "identity")], 8, 8, 4
rng = np.random.default_rng(4)
onet = ConvNet([ConvLayer(*l) for l in layers], W, H, time_dependent=True)
chain = pkg.TDConvChain(pkg.ConvChain(*[pkg.Conv(*l) for l in layers], width=W, height=H))
ps = glorot_uniform_conv_params(onet, rng, jitter=0.1)
x = rng.standard_normal((onet.state_dims, B)).astype(np.float32)
calls = []          # per call: the statistics one update would blend in, and t
def f(u, t):
    onet.running, onet.track = np.zeros(onet.nstate, np.float32), True
    onet.momentum = 1.0
    y = onet.f(u, ps, t)
    calls.append((onet.running.copy(), float(t)))
    return y
osol = orc.solve_tsit5(f, x, np.float32(0), np.float32(1), abstol=1e-3, reltol=1e-3, maxiters=1000)
def ema(seq):
    r = initial_conv_state(onet)
    for s in seq:
        r = np.float32(0.9) * r + np.float32(0.1) * s
    return r
stats = [c[0] for c in calls]
print("calls", len(calls), "times", [round(c[1], 4) for c in calls])
g = pkg.NeuralODE(chain, regularize="none", abstol=1e-3, reltol=1e-3, maxiters=1000)
sol, st2 = g(x, ps, g.initialstates(np.random.default_rng(5)))
r = st2["model"]["running"]
L1 = slice(0, 12)
print("plain", np.abs(ema(stats) - r)[L1].max(), np.abs(ema(stats) - r)[12:].max())
for k in range(1, 6):
    g = pkg.NeuralODE(chain, regularize="none", abstol=1e-3, reltol=1e-3, maxiters=k)
    sol, st2 = g(x, ps, g.initialstates(np.random.default_rng(5)))
    r = st2["model"]["running"]
    e = ema(stats[:3 + 6 * k])
    print("maxiters", k, sol.retcode, "nfe", st2["nfe"], "L1 mean", np.abs(e - r)[0:6].max(), "L1 var", np.abs(e - r)[6:12].max(),
          "L2", np.abs(e - r)[12:].max())
# standalone closure calls on the oracle's own inputs
inputs = []
def f2(u, t):
    inputs.append((u.copy(), float(t)))
    return onet.f(u, ps, t)
onet.track = False
orc.solve_tsit5(f2, x, np.float32(0), np.float32(1), abstol=1e-3, reltol=1e-3, maxiters=1)
mst = g.initialstates(np.random.default_rng(5))["model"]
for i, (u, t) in enumerate(inputs):
    _, mst = g.dynamics(u, ps, t, model_state=mst)
    e = ema(stats[:i + 1])
    print("standalone call", i, "t", round(t, 4), "L1", np.abs(e - mst["running"])[:12].max(), "L2", np.abs(e - mst["running"])[12:].max())
