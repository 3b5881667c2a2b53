"""Generates tests/golden/sde_*.npz from the CPU oracle (oracle/lrnde_sde_oracle.py).

The reference's NeuralDSDE tests (test/runtests.jl:340-430) assert only finiteness /
non-zeroness and its Wiener process is unseeded, so these fixtures pin the ORACLE restatement
(Philox noise): the CPU suite checks the oracle still reproduces them, the GPU suite checks
libLRNDE.so against them.

    python tests/golden/make_golden_sde.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as orc  # noqa: E402
from oracle import lrnde_sde_oracle as so  # noqa: E402

# name: (D, H, B, diffusion act, kwargs, controller overrides, weight scales, training)
CASES = {
    "sde_mnist_shape": (32, 64, 16, "identity", dict(regularize="unbiased", abstol=0.14, reltol=0.14, seed=0),
                        None, (1.0, 1.0), True),
    "sde_tight_tanh": (8, 16, 37, "tanh", dict(regularize="unbiased", abstol=0.02, reltol=0.02, seed=1),
                       None, (2.0, 1.5), True),
    "sde_biased": (8, 16, 40, "tanh", dict(regularize="biased", abstol=0.05, reltol=0.05, seed=2, save_start=False),
                   None, (2.0, 1.0), True),
    "sde_saveat": (6, 12, 33, "tanh", dict(regularize="unbiased", abstol=0.05, reltol=0.05, seed=4,
                                           saveat=[0.25, 0.5, 1.0]), None, (2.0, 1.0), True),
    "sde_eval": (6, 12, 33, "tanh", dict(regularize="unbiased", abstol=0.05, reltol=0.05, seed=4),
                 None, (2.0, 1.0), False),
    "sde_rejections": (2, 8, 1, "tanh", dict(regularize="none", abstol=0.05, reltol=0.05, seed=3),
                       dict(qmax=10.0), (3.0, 2.0), True),
}


def build(name):
    D, H, B, dact, kw, ctrl, (sf, sg), training = CASES[name]
    od = orc.MLP([orc.Dense(D, H, "tanh"), orc.Dense(H, D, "identity")], time_dependent=False)
    og = orc.MLP([orc.Dense(D, D, dact)], time_dependent=False)
    rng = np.random.default_rng(7)
    ps = np.concatenate([sf * orc.glorot_uniform_params(od, rng), sg * orc.glorot_uniform_params(og, rng)])
    ps = (ps + 0.02 * rng.standard_normal(ps.size)).astype(np.float32)
    x = rng.standard_normal((D, B)).astype(np.float32)
    return od, og, ps, x, rng


def run(name, dtype=np.float32):
    D, H, B, dact, kw, ctrl, _, training = CASES[name]
    od, og, ps, x, rng = build(name)
    saved = dict(so.SDE_CONSTS)
    if ctrl:
        so.SDE_CONSTS.update(ctrl)
    try:
        node = orc.NeuralDSDE(od, og, maxiters=10000, dtype=dtype, **kw)
        st = node.initialstates(np.random.default_rng(3))
        st["training"] = training
        out, st2, aux = node.forward(x.astype(dtype), ps.astype(dtype), st)
        cots = [(rng.standard_normal((D, B)) / B).astype(np.float32) for _ in out.u]
        d_reg = 0.5 if (kw["regularize"] != "none" and training) else 0.0
        dx, dps = node.backward(aux, [c.astype(dtype) for c in cots], d_reg, ps.astype(dtype))
    finally:
        so.SDE_CONSTS.clear(); so.SDE_CONSTS.update(saved)
    sol = aux["sol"]
    log = np.array([[float(l[0]), float(l[1]), float(l[2]), float(l[3])] for l in sol.log], np.float64)
    return dict(x=x, ps=ps, t=np.array(out.t, np.float64), u=np.stack(out.u).astype(np.float64),
                cots=np.stack(cots), d_reg=np.float64(d_reg), dx=dx.astype(np.float64), dps=dps.astype(np.float64),
                log=log, nfe=np.array([st2["nfe_drift"], st2["nfe_diffusion"]]), reg=np.float64(st2["reg_val"]),
                t1=np.float64(aux.get("t1", 0.0)), naccept=np.int64(len(sol.steps)), ndraws=np.int64(sol.ndraws))


if __name__ == "__main__":
    for name in CASES:
        g = run(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
        print(name, "steps", int(g["naccept"]), "attempts", len(g["log"]), "rej", int((g["log"][:, 3] == 0).sum()),
              "nfe", g["nfe"], "reg", float(g["reg"]), "|dps|", float(np.linalg.norm(g["dps"])))
