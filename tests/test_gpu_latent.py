"""GPU parity, latent-ODE encoder: Recurrence(LatentGRUCell) through the C ABI against the
numpy oracle (oracle/lrnde_latent_oracle.py) on the same seeded inputs, at the physionet shape
(37 features -> 75 rows, 49 time points, hidden 40, latent 50, batch 256: experiments/src/config.jl:31-34) and at ragged small shapes."""
import numpy as np
import pytest

import __graft_entry__ as entry
import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    entry.build()
    return entry.load_package()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _series(rng, in_dim, T, B, p_obs=0.2):
    """synthetic irregular series in the reference's layout: vcat(data .* mask, mask, dt) (construct.jl:40,59)"""
    mask = (rng.random((in_dim, T, B)) < p_obs).astype(np.float32)
    mask[:, rng.integers(0, T), :] = 0.0                       # a time point nobody observes ...
    data = rng.standard_normal((in_dim, T, B)).astype(np.float32) * mask
    dt = np.zeros((1, T, B), np.float32)                        # ... with dt = 0 keeps the carry (latent_ode.jl:40-43)
    dt[0, 1:, :] = np.diff(np.sort(rng.random((T, B)), axis=0), axis=0)
    return np.concatenate([data, mask, dt], axis=0)


@pytest.mark.parametrize("in_dim,H,L,T,B", [(37, 40, 50, 49, 256), (3, 5, 2, 4, 1), (6, 9, 5, 11, 77), (37, 40, 20, 100, 300)])
def test_gru_recurrence_matches_oracle(pkg, in_dim, H, L, T, B):
    rng = np.random.default_rng(B)
    cell = pkg.LatentGRUCell(in_dim, H, L)
    F = cell.features
    assert cell.nparams() == orc.gru_nparams(F, H, L)
    ps = (orc.gru_init(rng, F, H, L) + 0.05 * rng.standard_normal(cell.nparams())).astype(np.float32)
    x = _series(rng, in_dim, T, B)
    rec = pkg.Recurrence(cell)
    y, st = rec(x, ps)
    oy, carries = orc.gru_recurrence(ps, x, F, H, L)
    assert y.shape == (2 * L, B) and rel(y, oy) < 1e-5
    d_y = (rng.standard_normal((2 * L, B)) / B).astype(np.float32)
    d_ps = rec.backward(st, d_y)
    o_dps = orc.gru_recurrence_backward(ps, x, F, H, L, carries, d_y)
    assert rel(d_ps, o_dps) < 1e-4
    rec.free(st)


def test_gru_torch_tensors_and_errors(pkg):
    import torch
    rng = np.random.default_rng(3)
    cell = pkg.LatentGRUCell(4, 6, 3)
    ps = cell.initialparameters(rng)
    x = _series(rng, 4, 7, 19)
    rec = pkg.Recurrence(cell)
    y, st = rec(x, ps)
    yt, stt = rec(torch.from_numpy(x).cuda(), torch.from_numpy(ps).cuda())
    assert yt.cpu().numpy().tobytes() == np.ascontiguousarray(y).tobytes()
    d = np.ones_like(y) / 19
    assert rel(rec.backward(stt, torch.from_numpy(d).cuda()).cpu().numpy(), rec.backward(st, d)) < 1e-6
    with pytest.raises(ValueError):
        rec(x[:5], ps)
    with pytest.raises(ValueError):
        rec(x, ps[:-1])
    rec.free(st); rec.free(stt)


def test_small_layers_match_oracle(pkg):
    """rec_to_gen / gen_to_data (lrnde_mlp_*), ReparameterizeLayer and the latent-ODE loss against numpy."""
    from oracle.lrnde_latent_oracle import _net_fwd, _net_vjp
    from oracle import philox_normal
    rng = np.random.default_rng(0)
    chain = pkg.Chain(pkg.Dense(10, 6, "tanh"), pkg.Dense(6, 8))
    ps = (pkg.glorot_uniform(chain, rng) + 0.1 * rng.standard_normal(pkg.nparams(chain))).astype(np.float32)
    W1 = ps[:60].reshape((6, 10), order="F"); b1 = ps[60:66]; W2 = ps[66:114].reshape((8, 6), order="F"); b2 = ps[114:]
    layers = [(W1, b1, "tanh", 0, 60), (W2, b2, "identity", 66, 114)]
    x = rng.standard_normal((10, 5, 7)).astype(np.float32)                    # (in, T, B): 35 columns
    y = pkg.mlp_forward(chain, ps, x)
    oy = _net_fwd(layers, x.reshape(10, -1, order="F")).reshape((8, 5, 7), order="F")
    assert y.shape == (8, 5, 7) and rel(y, oy) < 1e-5
    d_y = rng.standard_normal(y.shape).astype(np.float32)
    d_x, d_ps = pkg.mlp_backward(chain, ps, x, d_y)
    odps = np.zeros_like(ps)
    odx = _net_vjp(layers, x.reshape(10, -1, order="F"), d_y.reshape(8, -1, order="F"), odps).reshape(x.shape, order="F")
    assert rel(d_x, odx) < 1e-5 and rel(d_ps, odps) < 1e-5
    # reparameterisation with the shared Philox stream
    L, B = 4, 9
    z = rng.standard_normal((2 * L, B)).astype(np.float32)
    rp = pkg.ReparameterizeLayer(seed=5)
    yr, st = rp(z, None, rp.initialstates(np.random.default_rng(0)))
    eps = philox_normal(5, 2, 0, L * B).reshape((L, B), order="F")
    oyr, mu, ls = orc.reparameterize(z, eps, True)
    assert rel(yr, oyr) < 1e-6 and np.array_equal(st["mu0"], z[:L]) and np.array_equal(st["logsigma2"], z[L:])
    dyr, dmu, dls = (rng.standard_normal((L, B)).astype(np.float32) for _ in range(3))
    dz = rp.backward(z, dyr, dmu, dls)
    assert rel(dz[:L], dyr + dmu) < 1e-6 and rel(dz[L:], dyr * eps * np.exp(z[L:] / 2) / 2 + dls) < 1e-5
    ye, _ = rp(z, None, dict(training=False))
    assert np.array_equal(np.asarray(ye), z[:L])                                # common.jl:73-77
    # loss
    F_, T_ = 5, 6
    mask = (rng.random((F_, T_, B)) < 0.4).astype(np.float32); mask[0, 0, :] = 1
    data = rng.standard_normal((F_, T_, B)).astype(np.float32) * mask
    pred = (data + 0.01 * rng.standard_normal((F_, T_, B))).astype(np.float32)
    w_kl = 0.3
    loss, nll, kl, d_pred, d_mu, d_ls = pkg.latent_loss(pred, data, mask, z[:L], z[L:], w_kl)
    ll = orc.log_likelihood_loss(pred * mask - data * mask, mask)
    okl = orc.kl_divergence(z[:L], z[L:])
    assert abs(loss - float(-(ll - w_kl * okl).mean())) < 1e-4 * abs(loss)
    assert abs(nll - float(-ll.mean())) < 1e-4 * abs(nll) and abs(kl - float(okl.mean())) < 1e-5
    msum = mask.sum(axis=(0, 1))
    assert rel(d_pred, (pred * mask - data * mask) * mask / 0.01 ** 2 / msum / B) < 1e-5
    p2, p3 = pred.copy(), pred.copy()                                     # central difference through the C ABI
    p2[0, 0, 3] += 1e-3; p3[0, 0, 3] -= 1e-3
    fd = (pkg.latent_loss(p2, data, mask, z[:L], z[L:], w_kl)[0] - pkg.latent_loss(p3, data, mask, z[:L], z[L:], w_kl)[0]) / 2e-3
    assert abs(fd - d_pred[0, 0, 3]) < 2e-2 * abs(d_pred[0, 0, 3]) + 2e-3
    assert rel(d_mu, w_kl * z[:L] / (L * B)) < 1e-5 and rel(d_ls, w_kl * (np.exp(z[L:]) - 1) / (2 * L * B)) < 1e-5


def test_physionet_latent_ode_end_to_end(pkg):
    """BASELINE configs[2] (physionet) end to end on the GPU against the oracle pipeline:
    Recurrence(LatentGRUCell(37, 40, 50)) -> rec_to_gen -> ReparameterizeLayer -> NeuralODE(8 x Dense(20<->40, tanh),
    saveat = 49 observation times, :unbiased local reg) -> diffeqsol_to_timeseries -> gen_to_data -> latent-ODE loss
    (experiments/src/construct.jl:36-70, 229-247; dims experiments/src/config.jl:31-34), forward and the whole
    reverse pass (loss -> gen_to_data -> adjoint with 49 jumps -> reparameterisation -> rec_to_gen -> BPTT)."""
    from oracle.lrnde_latent_oracle import _net_fwd, _net_vjp
    from oracle import philox_normal
    in_dim, H, Lg, Nn, T, B = 37, 40, 50, 20, 49, 256
    rng = np.random.default_rng(8)
    x = _series(rng, in_dim, T, B)
    data, mask = x[:in_dim], x[in_dim:2 * in_dim]
    ts = np.sort(np.concatenate([[0.0], rng.uniform(0.02, 1.0, T - 1)])).astype(np.float32)
    cell = pkg.LatentGRUCell(in_dim, H, Lg)
    F = cell.features
    ps_gru = (orc.gru_init(rng, F, H, Lg) * 1.0).astype(np.float32)
    r2g = pkg.Chain(pkg.Dense(2 * Lg, Lg, "tanh"), pkg.Dense(Lg, 2 * Nn))
    g2d = pkg.Chain(pkg.Dense(Nn, in_dim))
    ps_r2g, ps_g2d = pkg.glorot_uniform(r2g, rng), pkg.glorot_uniform(g2d, rng)
    dyn_layers = [(Nn, H, "tanh"), (H, Nn, "tanh")] * 4
    om = orc.MLP([orc.Dense(*l) for l in dyn_layers], time_dependent=False, input_act="tanh")
    ps_node = (orc.glorot_uniform_params(om, rng) * 2 + 0.05 * rng.standard_normal(om.nparams)).astype(np.float32)
    kw = dict(regularize="unbiased", abstol=1e-4, reltol=1e-4, maxiters=10000, saveat=list(ts))
    w_reg, w_kl, seed = 0.5, 0.1, 13

    # ---------------- GPU
    rec = pkg.Recurrence(cell)
    h, st_rec = rec(x, ps_gru)
    z = pkg.mlp_forward(r2g, ps_r2g, h)
    rp = pkg.ReparameterizeLayer(seed=seed)
    y0, st_rp = rp(z, None, dict(training=True))
    node = pkg.NeuralODE(pkg.Chain(*[pkg.Dense(*l) for l in dyn_layers], input_activation="tanh"), precision="tf32x3", **kw)
    sol, st_node = node(np.ascontiguousarray(y0), ps_node, node.initialstates(np.random.default_rng(2)))
    series = pkg.diffeqsol_to_timeseries(sol)                                   # (Nn, T, B)
    pred = pkg.mlp_forward(g2d, ps_g2d, series)
    loss, nll, kl, d_pred, d_mu, d_ls = pkg.latent_loss(pred, data, mask, st_rp["mu0"], st_rp["logsigma2"], w_kl)
    loss_total = loss + w_reg * float(st_node["reg_val"])
    d_series, d_g2d = pkg.mlp_backward(g2d, ps_g2d, series, d_pred)
    d_y0, d_node = node.backward(sol, [np.ascontiguousarray(d_series[:, i, :]) for i in range(T)], w_reg)
    d_z = rp.backward(z, d_y0, d_mu, d_ls)
    d_h, d_r2g = pkg.mlp_backward(r2g, ps_r2g, h, d_z)
    d_gru = rec.backward(st_rec, d_h)

    # ---------------- oracle
    def chain_layers(ps, dims):
        out, off = [], 0
        for (i, o, a) in dims:
            W = ps[off:off + o * i].reshape((o, i), order="F"); wo = off; off += o * i
            b = ps[off:off + o]; bo = off; off += o
            out.append((W, b, a, wo, bo))
        return out
    oh, carries = orc.gru_recurrence(ps_gru, x, F, H, Lg)
    Lr = chain_layers(ps_r2g, [(2 * Lg, Lg, "tanh"), (Lg, 2 * Nn, "identity")])
    Ld = chain_layers(ps_g2d, [(Nn, in_dim, "identity")])
    oz = _net_fwd(Lr, oh)
    eps = philox_normal(seed, 2, 0, Nn * B).reshape((Nn, B), order="F")
    oy0, omu, ols = orc.reparameterize(oz, eps, True)
    on = orc.NeuralODE(om, **kw)
    osol, ost, aux = on.forward(oy0, ps_node, on.initialstates(np.random.default_rng(2)))
    oseries = np.stack(osol.u, axis=1)
    opred = _net_fwd(Ld, oseries.reshape(Nn, -1, order="F")).reshape((in_dim, T, B), order="F")
    ll = orc.log_likelihood_loss(opred * mask - data * mask, mask)
    okl = orc.kl_divergence(omu, ols)
    oloss = float(-(ll - w_kl * okl).mean()) + w_reg * float(ost["reg_val"])
    msum = mask.sum(axis=(0, 1))
    o_dpred = ((opred * mask - data * mask) * mask / 0.01 ** 2 / msum / B).astype(np.float32)
    o_dg2d = np.zeros_like(ps_g2d)
    o_dseries = _net_vjp(Ld, oseries.reshape(Nn, -1, order="F"), o_dpred.reshape(in_dim, -1, order="F"), o_dg2d).reshape(oseries.shape, order="F")
    o_dy0, o_dnode = on.backward(aux, [o_dseries[:, i, :] for i in range(T)], w_reg, ps_node)
    o_dmu = w_kl * omu / (Nn * B); o_dls = w_kl * (np.exp(ols) - 1) / (2 * Nn * B)
    o_dz = np.concatenate([o_dy0 + o_dmu, o_dy0 * eps * np.exp(ols / 2) / 2 + o_dls], axis=0).astype(np.float32)
    o_dr2g = np.zeros_like(ps_r2g)
    o_dh = _net_vjp(Lr, oh, o_dz, o_dr2g)
    o_dgru = orc.gru_recurrence_backward(ps_gru, x, F, H, Lg, carries, o_dh)

    assert rel(h, oh) < 1e-5 and rel(z, oz) < 1e-5 and rel(y0, oy0) < 1e-5
    assert st_node["nfe"] == ost["nfe"] and len(sol.u) == T
    assert rel(series, oseries) < 1e-4 and rel(pred, opred) < 1e-4
    assert abs(loss_total / oloss - 1) < 1e-3
    assert rel(d_node, o_dnode) < 1e-3 and rel(d_g2d, o_dg2d) < 1e-3
    assert rel(d_y0, o_dy0) < 1e-3 and rel(d_r2g, o_dr2g) < 1e-3 and rel(d_gru, o_dgru) < 2e-3
    rec.free(st_rec); sol.free()
