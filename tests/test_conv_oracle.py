"""Self-validation of the conv-dynamics oracle (oracle/lrnde_conv_oracle.py).  CPU only."""
import numpy as np
import scipy.signal

import oracle as orc
from oracle.lrnde_conv_oracle import (ConvLayer, ConvNet, cifar10_node_core, glorot_uniform_conv_params,
                                      initial_conv_state)


def _small(dtype=np.float64, td=True, seed=0, B=3):
    rng = np.random.default_rng(seed)
    net = ConvNet([ConvLayer(2, 5, True, "gelu"), ConvLayer(5, 4, True, "gelu"), ConvLayer(4, 2)], 8, 4,
                  time_dependent=td)
    ps = glorot_uniform_conv_params(net, rng, dtype, jitter=0.2)
    u = rng.standard_normal((net.state_dims, B)).astype(dtype)
    return net, ps, u, rng


def test_parameter_count_of_the_cifar10_core():
    assert cifar10_node_core().nparams == 47560
    assert cifar10_node_core().state_dims == 32 * 32 * 8


def test_conv_is_a_true_convolution_with_zero_padding():
    """NNlib ``conv`` flips the kernel: the single-channel case equals scipy's convolve2d 'same'."""
    rng = np.random.default_rng(1)
    net = ConvNet([ConvLayer(1, 1)], 8, 4, time_dependent=False)
    w = rng.standard_normal(9)
    x = rng.standard_normal((4, 8))                        # [h, w]
    y = net.f(x.reshape(-1, 1), w, 0.0).reshape(4, 8)
    k = w.reshape((3, 3), order="F").T                     # w[kx, ky] -> k[ky, kx]
    np.testing.assert_allclose(y, scipy.signal.convolve2d(x, k, mode="same"), rtol=1e-12, atol=1e-12)


def test_time_channel_is_not_a_bias_at_the_border():
    """t * ones is concatenated BEFORE the zero padding (common.jl:19-33): border pixels see fewer time taps."""
    net = ConvNet([ConvLayer(1, 1)], 4, 4, time_dependent=True)
    ps = np.zeros(net.nparams)
    ps[9:] = 1.0                                           # only the time channel's taps
    y = net.f(np.zeros((16, 1)), ps, 2.0).reshape(4, 4)
    assert y[1, 1] == 18.0 and y[0, 0] == 8.0 and y[0, 1] == 12.0


def test_vjp_matches_finite_differences():
    net, ps, u, rng = _small()
    lam = rng.standard_normal(u.shape)
    a, dps = net.vjp(u, ps, 0.3, lam)
    eps = 1e-6
    for i in rng.choice(ps.size, 24, replace=False):
        p1, p2 = ps.copy(), ps.copy()
        p1[i] += eps; p2[i] -= eps
        fd = (np.sum(lam * net.f(u, p1, 0.3)) - np.sum(lam * net.f(u, p2, 0.3))) / (2 * eps)
        assert abs(dps[i] - fd) < 1e-6 * max(1.0, abs(fd)), (i, dps[i], fd)
    for i in rng.choice(u.size, 12, replace=False):
        u1, u2 = u.copy(), u.copy()
        u1.flat[i] += eps; u2.flat[i] -= eps
        fd = (np.sum(lam * net.f(u1, ps, 0.3)) - np.sum(lam * net.f(u2, ps, 0.3))) / (2 * eps)
        assert abs(a.flat[i] - fd) < 1e-6 * max(1.0, abs(fd))


def test_batchnorm_couples_the_batch():
    """Training-mode statistics are taken over the whole batch: changing one sample moves the others' outputs."""
    net, ps, u, _ = _small()
    y0 = net.f(u, ps, 0.1)
    u2 = u.copy(); u2[:, 0] *= 3.0
    y1 = net.f(u2, ps, 0.1)
    assert np.abs(y1[:, 1] - y0[:, 1]).max() > 1e-6


def test_neural_ode_layer_runs_on_the_conv_dynamics():
    """The reference's NeuralODE property test (test/runtests.jl:118-131) on the conv core: finite outputs,
    reg_val > 0, a gradient for every parameter block."""
    net, ps, u, rng = _small(np.float32, B=2)
    node = orc.NeuralODE(net, regularize="unbiased", abstol=1e-3, reltol=1e-3, maxiters=1000)
    st = node.initialstates(np.random.default_rng(0))
    out, st2, aux = node.forward(u, ps, st)
    assert all(np.isfinite(x).all() for x in out.u) and st2["reg_val"] > 0 and st2["nfe"] > 0
    d_us = [np.zeros_like(x) for x in out.u]
    d_us[-1] = rng.standard_normal(out.u[-1].shape).astype(np.float32)
    d_x, d_ps = node.backward(aux, d_us, np.float32(1.0), ps)
    assert np.isfinite(d_ps).all() and np.isfinite(d_x).all()
    for (wo, go), L in zip(net.offsets, net.layers):
        assert np.abs(d_ps[wo:go]).max() > 0


def test_running_statistics_follow_the_closure_and_testmode_uses_them():
    """Lux BatchNorm inside the ``dudt`` closure (neural_ode.jl:44-47): every training-mode call moves the running
    statistics by momentum 0.1 (unbiased variance); in testmode they replace the batch statistics and the
    samples decouple; the testmode pullback matches finite differences."""
    net, ps, u, rng = _small()
    net.running = initial_conv_state(net, np.float64)
    net.track = True
    z = net._conv(net._with_time(net._to_img(u, 2), 0.1), net.unpack(ps)[0][0])
    net.f(u, ps, 0.1)
    n = z.size // 5
    np.testing.assert_allclose(net.running[:5], 0.1 * z.mean(axis=(0, 2, 3)), rtol=1e-12)
    np.testing.assert_allclose(net.running[5:10], 0.9 + 0.1 * z.var(axis=(0, 2, 3)) * n / (n - 1), rtol=1e-12)
    net.track, net.testmode = False, True
    y0 = net.f(u, ps, 0.1)
    u2 = u.copy(); u2[:, 0] *= 3.0
    assert np.array_equal(net.f(u2, ps, 0.1)[:, 1:], y0[:, 1:])
    lam = rng.standard_normal(u.shape)
    a, dps = net.vjp(u, ps, 0.3, lam)
    eps = 1e-6
    for i in rng.choice(ps.size, 12, replace=False):
        p1, p2 = ps.copy(), ps.copy()
        p1[i] += eps; p2[i] -= eps
        fd = (np.sum(lam * net.f(u, p1, 0.3)) - np.sum(lam * net.f(u, p2, 0.3))) / (2 * eps)
        assert abs(dps[i] - fd) < 1e-6 * max(1.0, abs(fd))


def test_side_layers_match_finite_differences():
    """Conv with bias / activation, BatchNorm (training and testmode) and the AugmenterLayer: pullbacks vs central
    differences in Float64."""
    from oracle.lrnde_conv_oracle import augmenter, augmenter_vjp, batchnorm, batchnorm_vjp, conv2d, conv2d_vjp
    rng = np.random.default_rng(7)
    x = rng.standard_normal((8, 4, 3, 2))
    eps = 1e-6

    def check(fun, vjp_out, args, n=10):
        for k, (arr, grad) in enumerate(args):
            for i in rng.choice(arr.size, min(n, arr.size), replace=False):
                a1, a2 = arr.copy(), arr.copy()
                a1.flat[i] += eps; a2.flat[i] -= eps
                fd = (fun(*(a1 if j == k else a[0] for j, a in enumerate(args)))
                      - fun(*(a2 if j == k else a[0] for j, a in enumerate(args)))) / (2 * eps)
                assert abs(grad.flat[i] - fd) < 1e-6 * max(1.0, abs(fd)), (k, i, grad.flat[i], fd)

    ps = rng.standard_normal(9 * 3 * 5 + 5) * 0.3
    dy = rng.standard_normal((8, 4, 5, 2))
    d_x, d_ps = conv2d_vjp(x, ps, dy, "gelu")
    check(lambda xx, pp: np.sum(dy * conv2d(xx, pp, 5, "gelu")), None, [(x, d_x), (ps, d_ps)])
    assert augmenter(x, ps, 5).shape == (8, 4, 8, 2)
    da = rng.standard_normal((8, 4, 8, 2))
    d_x, d_ps = augmenter_vjp(x, ps, da)
    check(lambda xx, pp: np.sum(da * augmenter(xx, pp, 5)), None, [(x, d_x), (ps, d_ps)])
    bps = np.concatenate([1 + 0.2 * rng.standard_normal(3), 0.2 * rng.standard_normal(3)])
    dyb = rng.standard_normal(x.shape)
    d_x, d_ps = batchnorm_vjp(x, bps, dyb, "gelu")
    check(lambda xx, pp: np.sum(dyb * batchnorm(xx, pp, "gelu")[0]), None, [(x, d_x), (bps, d_ps)])
    run = np.concatenate([0.1 * rng.standard_normal(3), 0.5 + rng.random(3)])
    d_x, d_ps = batchnorm_vjp(x, bps, dyb, "identity", run, False)
    check(lambda xx, pp: np.sum(dyb * batchnorm(xx, pp, "identity", run, False)[0]), None, [(x, d_x), (bps, d_ps)])
    _, run2 = batchnorm(x, bps, "identity", np.concatenate([np.zeros(3), np.ones(3)]), True)
    n = x.size // 3
    np.testing.assert_allclose(run2[:3], 0.1 * x.mean(axis=(0, 1, 3)))
    np.testing.assert_allclose(run2[3:], 0.9 + 0.1 * x.var(axis=(0, 1, 3)) * n / (n - 1))
