"""GPU regression tests for state handling at the reference-facing API (round-1 advisor findings):
stale memoised factors after a same-size data replacement, cross-covariances of kernels that
slice / warp their inputs, bounds-cache staleness, argument-length checks before the C ABI,
and the row-strided dK/dtheta kernel beyond 65 535 rows."""
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")


def _data(n, dim, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((n, dim))
    y = np.sin(4 * x[:, 0]) + 0.3 * x[:, -1] + 0.05 * rng.standard_normal(n)
    return x, y


def test_replace_data_same_size_is_not_served_from_the_memo():
    """update_gp_data(append=False) with the same N, explicit noise and a custom (constant) prior mean: hps, m, V and N
    are unchanged, only x / y differ (gp.py:689-749).  State, LML, gradient and posterior must equal a fresh GP's."""
    from fvgp_b200 import GP
    n = 300
    x1, y1 = _data(n, 2, 1)
    x2, y2 = _data(n, 2, 2)
    noise = np.full(n, 1e-2)
    h = np.array([1.1, .3, .4])
    mean = lambda x, hps: np.full(len(x), 0.25)                      # noqa: E731
    gp = GP(x1, y1, init_hyperparameters=h, noise_variances=noise, prior_mean_function=mean)
    l1 = gp.log_likelihood(h)
    gp.update_gp_data(x2, y2, noise_variances_new=noise, append=False)
    fresh = GP(x2, y2, init_hyperparameters=h, noise_variances=noise, prior_mean_function=mean)
    assert gp.log_likelihood() == fresh.log_likelihood()
    assert gp.log_likelihood(h) == fresh.log_likelihood(h) != l1
    assert np.array_equal(gp.neg_log_likelihood_gradient(h), fresh.neg_log_likelihood_gradient(h))
    xp = np.random.default_rng(3).random((17, 2))
    assert np.array_equal(gp.posterior_mean(xp)["m(x)"], fresh.posterior_mean(xp)["m(x)"])
    # only x changes (same y, same everything else)
    gp.update_gp_data(x1, y2, noise_variances_new=noise, append=False)
    fresh = GP(x1, y2, init_hyperparameters=h, noise_variances=noise, prior_mean_function=mean)
    assert gp.log_likelihood(h) == fresh.log_likelihood(h)


def test_changed_args_reach_a_four_argument_kernel():
    from fvgp_b200 import GP
    from fvgp_b200 import kernels as K

    def kern(x1, x2, hps, args):
        return hps[0] * K.matern_kernel_diff1(K.get_distance_matrix(x1, x2), hps[1] * args["stretch"])
    x, y = _data(200, 2, 4)
    h = np.array([1.0, .4])
    gp = GP(x, y, init_hyperparameters=h, noise_variances=np.full(200, 1e-2), kernel_function=kern,
            args={"stretch": 1.0})
    a = gp.log_likelihood(h)
    gp.args = {"stretch": 2.0}
    b = gp.log_likelihood(h)
    gp.args["stretch"] = 1.0                                          # in-place edit of the dict
    c = gp.log_likelihood(h)
    assert a == c and a != b


def test_posterior_of_a_kernel_that_slices_its_inputs():
    """kernel built on x[:, :2] of 3-D inputs: training and posterior must use the same (sliced) point sets."""
    from fvgp_b200 import GP
    from fvgp_b200 import kernels as K
    from oracle import fvgp_oracle as orc

    def kern(x1, x2, hps):
        return hps[0] * K.squared_exponential_kernel(K.get_distance_matrix(x1[:, :2], x2[:, :2]), hps[1])
    n = 250
    x, y = _data(n, 3, 5)
    noise = np.full(n, 1e-2)
    h = np.array([1.2, .35])
    gp = GP(x, y, init_hyperparameters=h, noise_variances=noise, kernel_function=kern)
    xp = np.random.default_rng(6).random((40, 3))
    got_m = gp.posterior_mean(xp)["m(x)"]
    got_c = gp.posterior_covariance(xp)

    def d(a, b):
        return np.sqrt(((a[:, None, :2] - b[None, :, :2]) ** 2).sum(-1))
    Kxx = h[0] * np.exp(-d(x, x) ** 2 / (2 * h[1] ** 2)) + np.diag(noise)
    kxp = h[0] * np.exp(-d(x, xp) ** 2 / (2 * h[1] ** 2))
    kpp = h[0] * np.exp(-d(xp, xp) ** 2 / (2 * h[1] ** 2))
    m = np.mean(y)
    ref_m = m + kxp.T @ np.linalg.solve(Kxx, y - m)
    ref_S = kpp - kxp.T @ np.linalg.solve(Kxx, kxp)
    assert np.max(np.abs(got_m - ref_m)) <= 1e-8 * max(1.0, np.max(np.abs(ref_m)))
    assert np.max(np.abs(got_c["S"] - ref_S)) <= 1e-7
    del orc


def test_bounds_cache_sees_in_place_edits_and_recycled_arrays():
    from fvgp_b200 import kernels as K
    x = np.random.default_rng(0).random((64, 2))
    lo0, hi0 = K.point_bounds(x)
    x *= 100.0                                                        # in-place: same id, same address
    lo1, hi1 = K.point_bounds(x)
    assert np.allclose(hi1, 100.0 * hi0) and np.allclose(lo1, 100.0 * lo0)
    for k in range(50):                                               # arrays that die and whose addresses recycle
        z = np.full((64, 2), float(k))
        z[0, 0] = -k
        lo, hi = K.point_bounds(z)
        assert lo[0] == -k and hi[1] == k
        del z


def test_short_hyperparameters_raise_before_the_c_abi():
    from fvgp_b200 import GP, ops
    from fvgp_b200 import _lib as L
    x, y = _data(100, 3, 7)
    with pytest.raises(Exception):
        GP(x, y, init_hyperparameters=np.array([1.0, .3, .3]), noise_variances=np.full(100, 1e-2))
    xd = L.to_dev(x)
    with pytest.raises(ValueError):
        ops.kfill(L.K_MATERN32, xd, xd, 1.0, np.array([1.0, 2.0]))
    with pytest.raises(ValueError):
        ops.kgrad_dense_matern32(xd, xd, np.array([1.0, .3]))


def test_dense_kernel_gradient_beyond_65535_rows():
    """fvgp_kgrad_dense_matern32 strides rows over grid.y (the limit used to be n1 <= 65535)."""
    from fvgp_b200 import ops
    from fvgp_b200 import _lib as L
    rng = np.random.default_rng(8)
    n1, n2 = 70001, 24
    x1, x2 = rng.random((n1, 2)), rng.random((n2, 2))
    th = np.array([1.3, .4, .6])
    g = ops.kgrad_dense_matern32(L.to_dev(x1), L.to_dev(x2), th).cpu().numpy()
    rows = np.array([0, 1, 65534, 65535, 65536, n1 - 1])
    dx = x1[rows][:, None, :] - x2[None, :, :]
    d = np.sqrt(((dx / th[1:]) ** 2).sum(-1))
    a = np.sqrt(3.0) * d
    assert np.max(np.abs(g[0][rows] - (1 + a) * np.exp(-a))) <= 1e-13
    for i in range(2):
        ref = th[0] * 3.0 * dx[..., i] ** 2 / th[1 + i] ** 3 * np.exp(-a)
        assert np.max(np.abs(g[1 + i][rows] - ref)) <= 1e-12
