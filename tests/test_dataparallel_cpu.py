"""N > 1 design check on CPU (world_size 2, gloo): sharding the batch over ranks while
all-reducing (i) the sum of squares behind every step-size norm and (ii) the mu block of the
adjoint reproduces the single-process oracle run on the full batch -- same accepted / rejected
step sequence, same NFE, same states, same regulariser, and parameter gradients that add up.
This is exactly the set of exchanges libLRNDE.so performs through its peer-mapped mailboxes
(csrc/lrnde_kernels.cuh: lr_group_sum, mu_publish/gather/norm kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
import oracle.lrnde_oracle as lo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.default_rng(11)
    model = orc.MLP([orc.Dense(6, 9, "tanh"), orc.Dense(9, 6, "identity")], time_dependent=True)
    ps = (orc.glorot_uniform_params(model, rng) * 3 + 0.1 * rng.standard_normal(model.nparams)).astype(np.float32)
    x = rng.standard_normal((6, 8)).astype(np.float32)
    c = rng.standard_normal((6, 8)).astype(np.float32)
    kw = dict(regularize="unbiased", abstol=1e-4, reltol=1e-4)
    return model, ps, x, c, kw


def _allsum(v):
    t = torch.tensor([float(v)], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t.item())


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model, ps, x, c, kw = _problem()
    B = x.shape[1]
    lo_b, hi_b = rank * B // world, (rank + 1) * B // world
    xs, cs = x[:, lo_b:hi_b], c[:, lo_b:hi_b]
    P, n_loc = model.nparams, xs.size

    def rms_dp(v):
        T = v.dtype.type
        if v.ndim == 1 and v.size == n_loc + P:          # adjoint state [lambda_local ; mu_global]
            ss = _allsum(np.sum(v[:n_loc].astype(np.float64) ** 2)) + float(np.sum(v[n_loc:].astype(np.float64) ** 2))
            return np.sqrt(T(ss) / T(n_loc * world + P))
        return np.sqrt(T(_allsum(np.sum(v.astype(np.float64) ** 2))) / T(v.size * world))

    def reg_dp(utilde, uprev, u, abstol, reltol, dt):
        T = u.dtype.type
        r = lo.calculate_residuals(utilde, uprev, u, T(abstol), T(reltol))
        return np.sqrt(T(_allsum(np.sum(r.astype(np.float64) ** 2))) / T(u.size * world)) * T(dt)

    lo.rms = rms_dp
    lo.reg_error_estimate = reg_dp
    vjp0 = model.vjp

    def vjp_dp(u, ps_, t, lam):                          # mu is a batch sum: make it global
        a, dp = vjp0(u, ps_, t, lam)
        tt = torch.from_numpy(dp.astype(np.float64))
        dist.all_reduce(tt)
        return a, tt.numpy().astype(dp.dtype)
    model.vjp = vjp_dp

    node = orc.NeuralODE(model, **kw)
    st = node.initialstates(np.random.default_rng(3))
    sol, st2, aux = node.forward(xs, ps, st)
    d_x, d_ps = node.backward(aux, [None, cs], 0.0, ps)
    out[rank] = dict(u=sol.u[-1], nfe=st2["nfe"], reg=float(st2["reg_val"]), d_x=d_x, d_ps=d_ps,
                     log=[(float(t), float(dt), bool(a)) for (t, dt, e, a) in aux["sol"].step_log],
                     blog=[(float(t), float(dt), bool(a)) for (t, dt, e, a) in aux["bsol"].step_log])
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_batch_reproduces_single_process():
    world = 2
    model, ps, x, c, kw = _problem()
    node = orc.NeuralODE(model, **kw)
    st = node.initialstates(np.random.default_rng(3))
    sol, st2, aux = node.forward(x, ps, st)
    d_x, d_ps = node.backward(aux, [None, c], 0.0, ps)
    ref_log = [(float(t), float(dt), bool(a)) for (t, dt, e, a) in aux["sol"].step_log]
    ref_blog = [(float(t), float(dt), bool(a)) for (t, dt, e, a) in aux["bsol"].step_log]

    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    B = x.shape[1]
    for r in range(world):
        o = out[r]
        lo_b, hi_b = r * B // world, (r + 1) * B // world
        assert o["nfe"] == st2["nfe"]
        assert [a for (_, _, a) in o["log"]] == [a for (_, _, a) in ref_log]
        assert [a for (_, _, a) in o["blog"]] == [a for (_, _, a) in ref_blog]
        np.testing.assert_allclose([t for (t, _, _) in o["log"]], [t for (t, _, _) in ref_log], rtol=1e-4, atol=1e-6)
        # the adjoint's first error estimates sit at the Float32 noise floor (different summation
        # order of the sharded sums): same decisions, step times within a few 1e-3
        np.testing.assert_allclose([t for (t, _, _) in o["blog"]], [t for (t, _, _) in ref_blog], rtol=0.05, atol=0.02)
        np.testing.assert_allclose(o["u"], sol.u[-1][:, lo_b:hi_b], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(o["reg"], float(st2["reg_val"]), rtol=2e-3)
        np.testing.assert_allclose(o["d_x"], d_x[:, lo_b:hi_b], rtol=1e-3, atol=1e-5)
        # mu was made global inside the exchange: every rank already holds the full gradient
        np.testing.assert_allclose(o["d_ps"], d_ps, rtol=1e-3, atol=1e-5)
    # both ranks took bit-identical decisions
    assert out[0]["log"] == out[1]["log"] and out[0]["blog"] == out[1]["blog"]
