import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
import oracle as orc
pkg = entry.load_package()
def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
def case(name, D, H, B, dact, mode, tol, seed, controller=None, saveat=None, save_start=None, scale=1.0, dscale=1.0, training=True):
    dl = [(D, H, "tanh"), (H, D, "identity")]; gl = [(D, D, dact)]
    od = orc.MLP([orc.Dense(*l) for l in dl], time_dependent=False); og = orc.MLP([orc.Dense(*l) for l in gl], time_dependent=False)
    cd = pkg.Chain(*[pkg.Dense(*l) for l in dl]); cg = pkg.Chain(*[pkg.Dense(*l) for l in gl])
    rng = np.random.default_rng(7)
    ps = np.concatenate([scale * orc.glorot_uniform_params(od, rng), dscale * orc.glorot_uniform_params(og, rng)]).astype(np.float32)
    ps += (0.02 * rng.standard_normal(ps.size)).astype(np.float32)
    x = rng.standard_normal((D, B)).astype(np.float32)
    kw = dict(regularize=mode, abstol=tol, reltol=tol, maxiters=10000, seed=seed)
    if saveat is not None: kw["saveat"] = saveat
    if save_start is not None: kw["save_start"] = save_start
    on = orc.NeuralDSDE(od, og, **kw)
    if controller:
        import oracle.lrnde_sde_oracle as so
        saved = dict(so.SDE_CONSTS); so.SDE_CONSTS.update(controller)
    gn = pkg.NeuralDSDE(cd, cg, controller=controller, **kw)
    st_o = on.initialstates(np.random.default_rng(3)); st_g = gn.initialstates(np.random.default_rng(3))
    st_o["training"] = training; st_g["training"] = training
    out, ost2, aux = on.forward(x, ps, st_o)
    sol, gst2 = gn(x, ps, st_g, keep_tape=True)
    osol = aux["sol"]
    t, dt, ee, acc = sol.step_log()
    olog = osol.log
    print(f"== {name}: oracle steps {len(osol.steps)} attempts {len(olog)} rej {sum(1 for l in olog if not l[3])} ret {osol.retcode} | gpu acc {sol.stats.naccept} rej {sol.stats.nreject} ret {sol.retcode} S-launches {sol.stats.gpu_launches}")
    n = min(len(olog), len(t))
    same = all(bool(olog[i][3]) == bool(acc[i]) for i in range(n)) and len(olog) == len(t)
    print("   decisions identical:", same, " max rel dt diff", max(abs(float(olog[i][1]) - float(dt[i])) / float(olog[i][1]) for i in range(n)),
          " max rel EEst diff", max(abs(float(olog[i][2]) - float(ee[i])) / max(float(olog[i][2]), 1e-30) for i in range(n)))
    print("   nfe", gst2["nfe_drift"], ost2["nfe_drift"], gst2["nfe_diffusion"], ost2["nfe_diffusion"], " len u", len(sol.u), len(out.u))
    print("   t match", np.allclose(np.array(sol.t, np.float64), np.array(out.t, np.float64), rtol=1e-5), " state rel", [f"{rel(a, b):.2e}" for a, b in list(zip(sol.u, out.u))[:4]], f"last {rel(sol.u[-1], out.u[-1]):.2e}")
    if mode != "none" and training:
        print("   reg", float(gst2["reg_val"]), float(ost2["reg_val"]), " t1", sol.stats.t1_used, aux["t1"], " dt_reg", sol.stats.dt_reg, aux["dt_reg"])
    cots = [(rng.standard_normal((D, B)) / B).astype(np.float32) for _ in sol.u]
    dreg = 0.5 if (mode != "none" and training) else 0.0
    dx, dps = gn.backward(sol, cots, dreg)
    odx, odps = on.backward(aux, cots, dreg, ps)
    nf = od.nparams
    print(f"   grad rel: d_x {rel(dx, odx):.2e} d_drift {rel(dps[:nf], odps[:nf]):.2e} d_diffusion {rel(dps[nf:], odps[nf:]):.2e}")
    if dreg:
        dx0, dps0 = gn.backward(sol, [None] * len(sol.u), 1.0)
        _, odps0 = on.backward(aux, [None] * len(out.u), 1.0, ps)
        print(f"   reg-only grad rel {rel(dps0, odps0):.2e}  |g| {np.linalg.norm(odps0):.3e}  dx0 {np.abs(dx0).max():.1e}")
    sol.free()
    if controller:
        so.SDE_CONSTS.clear(); so.SDE_CONSTS.update(saved)
case("mnist_sde shape unbiased", 32, 64, 128, "identity", "unbiased", 0.14, 0)
case("tight tol unbiased tanh-diffusion", 8, 16, 37, "tanh", "unbiased", 0.02, 1, scale=2.0, dscale=1.5)
case("biased all steps", 8, 16, 40, "tanh", "biased", 0.05, 2, scale=2.0, save_start=False)
case("saveat unbiased", 6, 12, 33, "tanh", "unbiased", 0.05, 4, saveat=[0.25, 0.5, 1.0], scale=2.0)
case("eval mode", 6, 12, 33, "tanh", "unbiased", 0.05, 4, training=False)
for sd in range(6):
    case(f"rejections qmax=10 seed {sd}", 2, 8, 1, "tanh", "none", 0.05, sd, controller=dict(qmax=10.0), scale=3.0, dscale=2.0)
