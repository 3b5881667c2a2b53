"""GPU parity, latent-ODE encoder: Recurrence(LatentGRUCell) through the C ABI against the
numpy oracle (oracle/lrnde_latent_oracle.py) on the same seeded inputs, at the physionet shape
(37 features -> 75 rows, 49 time points, hidden 40, latent 20, batch 256) and at ragged small shapes."""
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


@pytest.mark.parametrize("in_dim,H,L,T,B", [(37, 40, 20, 49, 256), (3, 5, 2, 4, 1), (6, 9, 5, 11, 77), (37, 40, 20, 100, 300)])
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
