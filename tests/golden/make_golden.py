"""Generates tests/golden/*.npz from the CPU oracle (oracle/lrnde_oracle.py).

The reference (Julia) cannot run in this image and holds no golden vectors of its own
(test/runtests.jl asserts only finiteness), so these fixtures pin the ORACLE: the CPU suite
checks the oracle still reproduces them, the GPU suite checks libLRNDE.so against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as orc  # noqa: E402

CASES = {
    # name: (layers [(in,out,act)...], td, input_act, B, kwargs, seed, d_reg)
    "tiny_td_gelu": ([(2, 4, "gelu"), (4, 2, "identity")], True, None, 3,
                     dict(regularize="unbiased"), 0, 2.5),
    "tiny_plain_biased": ([(2, 4, "gelu"), (4, 2, "identity")], False, None, 1,
                          dict(regularize="biased"), 1, 1.0),
    "mid_tanh_stiff": ([(16, 12, "tanh"), (12, 16, "identity")], True, None, 8,
                       dict(regularize="unbiased", regularize_type="stiffness_estimate",
                            abstol=1e-6, reltol=1e-6), 2, 0.7),
    "latent_saveat": ([(20, 40, "tanh"), (40, 20, "tanh"), (20, 40, "tanh"), (40, 20, "tanh")],
                      False, "tanh", 16,
                      dict(regularize="unbiased", abstol=1e-6, reltol=1e-6,
                           saveat=[0.0, 0.13, 0.37, 0.5, 0.81, 1.0]), 3, 0.3),
    "mnist_b16": ([(784, 100, "tanh"), (100, 784, "identity")], True, None, 16,
                  dict(regularize="unbiased", abstol=1e-6, reltol=1e-6, maxiters=10000,
                       save_start=False), 4, 2.5),
    "eval_mode": ([(5, 7, "sigmoid"), (7, 5, "relu")], True, None, 6,
                  dict(regularize="unbiased", abstol=1e-7, reltol=1e-5), 5, 0.0),
}


def build(name):
    layers, td, input_act, B, kw, seed, d_reg = CASES[name]
    model = orc.MLP([orc.Dense(*l) for l in layers], time_dependent=td, input_act=input_act)
    rng = np.random.default_rng(seed)
    ps = orc.glorot_uniform_params(model, rng)
    if layers[0][0] < 100:
        ps = ps + (0.1 * rng.standard_normal(ps.size)).astype(np.float32)   # non-zero biases
    D = layers[0][0]
    x = (rng.random((D, B), dtype=np.float32) if D == 784
         else rng.standard_normal((D, B)).astype(np.float32))
    node = orc.NeuralODE(model, **kw)
    st = node.initialstates(np.random.default_rng(seed + 100))
    if name == "eval_mode":
        st["training"] = False
    sol, st2, aux = node.forward(x, ps, st)
    cots = [rng.standard_normal((D, B)).astype(np.float32) for _ in sol.u]
    if name in ("tiny_td_gelu", "mnist_b16"):
        cots[0] = None            # the loss reads sol.u[end] only
    d_x, d_ps = node.backward(aux, cots, d_reg, ps)
    s = aux["sol"]
    log = np.array([(t, dt, e, a) for (t, dt, e, a) in s.step_log], dtype=np.float64)
    blog = np.array([(t, dt, e, a) for (t, dt, e, a) in aux["bsol"].step_log], dtype=np.float64)
    out = dict(
        x=x, ps=ps, t=np.array(sol.t, np.float32), u=np.stack(sol.u, 0),
        cot=np.stack([np.zeros((D, B), np.float32) if c is None else c for c in cots], 0),
        cot_mask=np.array([c is not None for c in cots]),
        d_reg=np.float32(d_reg), d_x=d_x.astype(np.float32), nfe=np.int64(st2["nfe"]),
        reg_val=np.float32(st2["reg_val"]), naccept=np.int64(s.naccept), nreject=np.int64(s.nreject),
        step_log=log, bwd_step_log=blog, nf_bwd=np.int64(aux["bsol"].nf),
        t1=np.float32(aux.get("t1", 0.0)), dt_reg=np.float32(aux.get("dt_reg", 0.0)),
    )
    if d_ps.size > 20000:       # keep the fixture small: strided sample + norm
        out["d_ps_stride"] = np.int64(53)
        out["d_ps"] = d_ps[::53].astype(np.float32)
        out["d_ps_norm"] = np.float64(np.linalg.norm(d_ps.astype(np.float64)))
    else:
        out["d_ps_stride"] = np.int64(1)
        out["d_ps"] = d_ps.astype(np.float32)
        out["d_ps_norm"] = np.float64(np.linalg.norm(d_ps.astype(np.float64)))
    return out


if __name__ == "__main__":
    for name in CASES:
        out = build(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "nfe", out["nfe"], "reg", out["reg_val"], "steps", out["naccept"], out["nreject"],
              "bwd attempts", len(out["bwd_step_log"]))
