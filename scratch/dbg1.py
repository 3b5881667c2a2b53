import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
import oracle as orc
from oracle.lrnde_oracle import tsit5_stages
pkg = entry.load_package()
np.set_printoptions(precision=7, linewidth=200)
for td in (False, True):
    layers = [(2, 4, "gelu"), (4, 2, "identity")]
    om = orc.MLP([orc.Dense(*l) for l in layers], time_dependent=td)
    rng = np.random.default_rng(1)
    ps = orc.glorot_uniform_params(om, rng) + (0.1 * rng.standard_normal(om.nparams)).astype(np.float32)
    x = rng.standard_normal((2, 3)).astype(np.float32)
    c = pkg.Chain(*[pkg.Dense(*l) for l in layers])
    c = pkg.TDChain(c) if td else c
    for maxit in (1, 2, 1000):
        node = pkg.NeuralODE(c, regularize="none", precision="fp32", maxiters=maxit)
        st = node.initialstates(np.random.default_rng(0))
        sol, st2 = node(x, ps, st, keep_tape=True)
        t, dt, ee, acc = sol.step_log(0)
        f = lambda u, tt: om.f(u, ps, tt)
        osol = orc.solve_tsit5(f, x, 0.0, 1.0, abstol=1e-6, reltol=1e-3, maxiters=maxit)
        print("td", td, "maxit", maxit, sol.retcode)
        print(" gpu t ", t, "\n gpu dt", dt, "\n gpu ee", ee)
        print(" orc t ", np.array([s[0] for s in osol.step_log]), "\n orc dt", np.array([s[1] for s in osol.step_log]), "\n orc ee", np.array([s[2] for s in osol.step_log]))
        print(" u gpu", np.asarray(sol.u[-1]).ravel(), "\n u orc", osol.us[-1].ravel())
