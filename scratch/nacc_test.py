import os, sys, subprocess
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import os, sys, numpy as np, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath("%s"))))
import __graft_entry__ as entry, oracle as orc
pkg = entry.load_package(); lib = pkg.lib()
layers=[(784,100,"tanh"),(100,784,"identity")]
om = orc.MLP([orc.Dense(*l) for l in layers], time_dependent=True)
rng=np.random.default_rng(0)
ps = orc.glorot_uniform_params(om, rng) + (0.05*rng.standard_normal(om.nparams)).astype(np.float32)
x = (3*rng.standard_normal((784,256))).astype(np.float32)
layer = pkg.NeuralODE(pkg.TDChain(pkg.Chain(*[pkg.Dense(*l) for l in layers])), precision="tf32x3")
got = layer.dynamics(x, ps, 0.37)
want64 = om.f(x.astype(np.float64), ps.astype(np.float64), 0.37)
want32 = om.f(x, ps, np.float32(0.37))
e = lambda a,b: float(np.abs(a-b).max()/np.abs(b).max())
print("nacc", os.environ.get("LRNDE_NACC"), "err vs f64: gpu %%.2e  numpy-f32 %%.2e" %% (e(got,want64), e(want32,want64)))
''' % os.path.join(here, "x")
for n in ("8", "4", "2", "1"):
    env = dict(os.environ, LRNDE_NACC=n)
    print(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip())
    env["LRNDE_PROFILE_LAYERS"] = "1"
    for ns in ("0", "5"):
        env["LRNDE_PROFILE_NSRC"] = ns
        out = subprocess.run([sys.executable, os.path.join(here, "prof_feval.py"), "8192", "50"], env=env, capture_output=True, text=True)
        print("   layer1 nsrc", ns, out.stdout.strip().split("TFLOP")[0])
