"""Running statistics / f of the tensor-core conv engine against the Float64 oracle and the SIMT engine."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
from oracle.lrnde_conv_oracle import ConvLayer, ConvNet, glorot_uniform_conv_params, initial_conv_state
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / np.linalg.norm(np.asarray(b, np.float64)))
cl, W, H = [(8, 16, True, "gelu"), (16, 16, True, "gelu"), (16, 8, False, "identity")], 32, 4
for B in (2, 4):
    onet = ConvNet([ConvLayer(*l) for l in cl], W, H, time_dependent=True)
    chain = pkg.TDConvChain(pkg.ConvChain(*[pkg.Conv(*l) for l in cl], width=W, height=H))
    rng = np.random.default_rng(11)
    ps = glorot_uniform_conv_params(onet, rng, jitter=0.2)
    x = rng.standard_normal((onet.state_dims, B)).astype(np.float32)
    node = pkg.NeuralODE(chain)
    st = node.initialstates(np.random.default_rng(0))
    onet.running, onet.track = initial_conv_state(onet, np.float64), True
    want = onet.f(x.astype(np.float64), ps.astype(np.float64), 0.2)
    for eng in ("1", "0"):
        os.environ["LRNDE_CONV_TC"] = eng
        du, mst = node.dynamics(x, ps, 0.2, model_state=st["model"])
        r = np.asarray(mst["running"], np.float64)
        print(f"B {B} TC={eng}: f {rel(du, want):.2e} running {rel(r, onet.running):.2e} per-entry max rel {np.max(np.abs(r - onet.running) / np.abs(onet.running)):.2e}")
        print("   running (gpu)   ", r[:4], r[16:20])
        print("   running (oracle)", onet.running[:4], onet.running[16:20])

# whole solve: running statistics / u / reg of the two engines on one GPU
print("whole solve")
B = 4
onet = ConvNet([ConvLayer(*l) for l in cl], W, H, time_dependent=True)
chain = pkg.TDConvChain(pkg.ConvChain(*[pkg.Conv(*l) for l in cl], width=W, height=H))
rng = np.random.default_rng(11)
ps = glorot_uniform_conv_params(onet, rng, jitter=0.2)
x = rng.standard_normal((onet.state_dims, B)).astype(np.float32)
out = {}
for eng in ("1", "0", "1"):
    os.environ["LRNDE_CONV_TC"] = eng
    node = pkg.NeuralODE(chain, regularize="unbiased", abstol=1e-3, reltol=1e-3, maxiters=1000, save_start=False)
    sol, st2 = node(x, ps, node.initialstates(np.random.default_rng(5)))
    t, dt, ee, acc = sol.step_log(0)
    out[eng] = (np.asarray(sol.u[-1]).copy(), np.asarray(st2["model"]["running"]).copy(), st2["nfe"], acc.copy())
    print("TC", eng, "nfe", st2["nfe"], "acc", acc, "running[6]", out[eng][1][6], "running[23]", out[eng][1][23])
    sol.free()
print("u TC vs SIMT", rel(out["1"][0], out["0"][0]), "running TC vs SIMT", rel(out["1"][1], out["0"][1]))
