"""Pin the numpy oracle against fixtures produced by the unmodified reference."""
import numpy as np
import scipy.sparse as sp

from oracle import fvgp_oracle as orc


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0


def test_dense_kernels(golden):
    g = golden("dense_kernels")
    x1, x2, h = g["x1"], g["x2"], g["hps"]
    assert rel(orc.default_kernel(x1, x2, h), g["default_12"]) <= 1e-15
    assert rel(orc.default_kernel(x1, x1, h), g["default_11"]) <= 1e-15
    assert np.max(np.abs(orc.default_kernel_gradient(x1, x1, h) - g["default_grad_11"])) <= 1e-15
    assert np.max(np.abs(orc.default_kernel_gradient(x1, x2, h) - g["default_grad_12"])) <= 1e-15
    assert np.array_equal(orc.distance_matrix(x1, x2), g["d_iso"])
    assert np.array_equal(orc.anisotropic_distance_matrix(x1, x2, h[1:]), g["d_ani"])
    for nm in ("se", "exp", "matern32", "matern52"):
        assert rel(orc.RADIAL[nm](g["d_iso"], 0.37), g[nm + "_iso"]) <= 1e-15
        assert rel(orc.RADIAL[nm](g["d_ani"], 1.3), g[nm + "_ani"]) <= 1e-15
    assert rel(orc.default_kernel(g["c1_x"], g["c1_x"], g["c1_hps"]), g["c1_K"]) <= 1e-15


def test_robust_and_wendland_kernels(golden):
    g = golden("robust_kernels")
    x1, x2, h, phi = g["x1"], g["x2"], g["hps"], float(g["phi"])
    d_iso, d_ani = orc.distance_matrix(x1, x2), orc.anisotropic_distance_matrix(x1, x2, h[1:])
    for nm in ("se", "exp", "matern32", "matern52"):
        assert rel(orc.RADIAL_ROBUST[nm](d_iso, phi), g[nm + "_robust_iso"]) <= 1e-15
        assert rel(orc.RADIAL_ROBUST[nm](d_ani, phi), g[nm + "_robust_ani"]) <= 1e-15
    assert np.array_equal(orc.wendland(d_ani), g["wendland_kernel_ani"])
    assert np.array_equal(orc.wendland_anisotropic(x1, x2, h), g["wendland_anisotropic_12"])
    assert np.array_equal(orc.wendland_anisotropic(x1, x1, h), g["wendland_anisotropic_11"])
    # the block kernel of gp2Scale (a16) and the support-aware sparse variant (a17) against the dense Wendland
    assert np.array_equal(orc.wendland_block(x1, x2, h), g["wendland_block_12"])
    assert np.array_equal(g["wendland_block_12"] != 0, g["wendland_sparse_12"] != 0)
    assert np.max(np.abs(g["wendland_sparse_12"] - g["wendland_block_12"])) <= 1e-12


def test_dense_lml_and_gradient(golden):
    for tag in ("c1", "c2"):
        g = golden("dense_lml_" + tag)
        x, y, nz = g["x"], g["y"], g["noise"]
        for hk in ("h0", "h1"):
            assert abs(orc.dense_log_likelihood(x, y, g[hk], nz) / g["lml_" + hk] - 1) <= 1e-11
            for eco in (False, True):
                gr = orc.dense_neg_log_likelihood_gradient(x, y, g[hk], nz, economical=eco)
                assert rel(gr, g["grad_" + hk]) <= 1e-8, (tag, hk, eco, gr, g["grad_" + hk])
        assert abs(g["lml_h0_state"] / g["lml_h0"] - 1) <= 1e-12
        assert abs(orc.dense_log_likelihood(x[:200], y[:200], g["h1"]) / g["lml_default_noise"] - 1) <= 1e-11

        def se(a, b, h):
            return h[0] * orc.squared_exponential(orc.distance_matrix(a, b), h[1])
        assert abs(orc.dense_log_likelihood(x[:300], y[:300], g["se_hps"] * 1.1, nz[:300], kernel=se)
                   / g["lml_se"] - 1) <= 1e-10
        mean, S = orc.posterior_mean_cov(x, y, g["h0"], nz, g["x_pred"])
        assert np.allclose(mean, g["post_mean"], rtol=1e-9, atol=1e-11)
        assert np.allclose(S, g["post_S"], rtol=1e-7, atol=1e-10)


def test_multitask_transform_and_lml(golden):
    g = golden("multitask")
    xi, yf, vf = orc.fvgp_transform(g["x"], g["y"], g["noise"])
    assert np.array_equal(xi, g["x_index"])
    assert np.array_equal(yf, g["y_flat"][:, 0])
    assert np.array_equal(vf, g["v_flat"])
    assert abs(orc.dense_log_likelihood(xi, yf, g["h1"], vf) / g["lml"] - 1) <= 1e-11
    assert rel(orc.dense_neg_log_likelihood_gradient(xi, yf, g["h1"], vf, economical=True), g["grad"]) <= 1e-8


def test_wendland_blocks_bit_exact(golden):
    g = golden("gp2scale_blocks")
    assert np.array_equal(orc.wendland_block(g["t1711_x1"], g["t1711_x2"], g["t1711_hps"]), g["t1711_dense"])
    d = orc.wendland_block(g["t1724_x"], g["t1724_x"], g["t1724_hps"])
    assert np.array_equal(d, g["t1724_dense"])
    assert np.all(np.diag(d) == g["t1724_hps"][0])                       # tests/test_fvgp.py:1724
    x1, x2, h = g["t3027_x1"], g["t3027_x2"], g["t3027_hps"]
    Ks = orc.gp2scale_covariance(x1, x1, h, batch=10, symmetric=True)
    Kr = orc.gp2scale_covariance(x1, x2, h, batch=10, symmetric=False)
    for K, p in ((Ks, "t3027_sym_"), (Kr, "t3027_rect_")):
        assert K.indices.dtype == np.int32 and K.has_sorted_indices
        assert np.array_equal(K.indptr, g[p + "indptr"])
        assert np.array_equal(K.indices, g[p + "indices"])
        assert np.array_equal(K.data, g[p + "data"])
    assert (Ks != Ks.T).nnz == 0
    # disjoint blocks -> empty (tests/test_fvgp.py:1737)
    far = orc.gp2scale_covariance(np.zeros((5, 2)), np.ones((4, 2)) * 10, h, batch=3, symmetric=False)
    assert far.nnz == 0 and far.shape == (5, 4)


def test_gp2scale_lml(golden):
    for tag, batch in (("t3152", 25), ("c4small", 1000)):
        g = golden("gp2scale_" + tag)
        x, y, nz = g["x"], g["y"], g["noise"]
        for pre, hk in (("", "h0"), ("h1_", "h1")):
            K = orc.gp2scale_covariance(x, x, g[hk], batch=batch, symmetric=True)
            assert np.array_equal(K.indptr, g[pre + "indptr"])
            assert np.array_equal(K.indices, g[pre + "indices"])
            if pre + "data" in g:
                assert np.array_equal(K.data, g[pre + "data"])
            else:
                assert np.array_equal(K.data[:2000], g[pre + "data_head"])
                assert abs(K.data.sum() / g[pre + "data_sum"] - 1) < 1e-13
            # batch size must not change the matrix (tests/test_fvgp.py:3152)
            K2 = orc.gp2scale_covariance(x, x, g[hk], batch=batch * 3 + 1, symmetric=True)
            assert np.array_equal(K2.indices, K.indices) and np.array_equal(K2.data, K.data)
            assert abs(orc.gp2scale_log_likelihood(x, y, g[hk], nz, batch=batch) / g["lml_" + hk] - 1) <= 1e-10
        # CG at tight tolerance reproduces the exact solve
        KV = orc.add_kv(orc.gp2scale_covariance(x, x, g["h0"], batch=batch, symmetric=True), nz)
        ym = (y - y.mean())[:, None]
        sol, iters = orc.sparse_cg(KV, ym, rtol=1e-12)
        assert np.allclose(sol, g["KVinvY_h0"], rtol=1e-7, atol=1e-9) and iters[0] > 0
        assert sp.issparse(KV)


def test_large_n_oracle_variants_equal_the_pinned_ones(golden):
    """The memory-lean variants bench.py / the -m gpu suite use at N = 8 000 ... 50 000 (in-place LAPACK, blocked dK,
    threaded block assembly) against the reference's own outputs and the plain oracle functions."""
    for tag in ("c1", "c2"):
        g = golden("dense_lml_" + tag)
        x, y, nz = g["x"], g["y"], g["noise"]
        for hk in ("h0", "h1"):
            assert abs(orc.dense_log_likelihood_blocked(x, y, g[hk], nz, block=97) / g["lml_" + hk] - 1) <= 1e-11
            # the N^2 >= 2^31 branch (one level of the 2 x 2 block factorisation), forced at small N
            assert abs(orc.dense_log_likelihood_blocked(x, y, g[hk], nz, block=97, split=True) / g["lml_" + hk] - 1) <= 1e-11
            lml, gr = orc.dense_neg_log_likelihood_gradient_blocked(x, y, g[hk], nz, block=53, threads=3)
            assert abs(lml / g["lml_" + hk] - 1) <= 1e-11
            assert rel(gr, g["grad_" + hk]) <= 1e-8, (tag, hk, gr, g["grad_" + hk])
    rng = np.random.default_rng(7)
    x = rng.random((900, 3))
    th = np.array([1.3, .11, .12, .1])
    a = orc.gp2scale_covariance(x, x, th, batch=128, symmetric=True)
    b = orc.gp2scale_covariance(x, x, th, batch=128, symmetric=True, threads=4)
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)
