"""GPU parity tests: the CUDA path, called through the reference-facing API (GP / fvGP / kernels) and the
C ABI, against the golden fixtures made by the unmodified reference and against the oracle.

Tolerances (BASELINE.json north_star): K entries <= 1e-12 relative, LML and gradient <= 1e-8
relative, gp2Scale sparsity pattern bit-exact."""
import os
import pickle
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


@pytest.fixture(scope="module")
def fv():
    import torch
    assert torch.cuda.is_available()
    import fvgp_b200
    from fvgp_b200 import _lib
    _lib.load()
    return fvgp_b200


def test_kernel_probe_sections_pass(fv):
    """Every low-level kernel check of tests/gpu_probe.py (shapes, edge cases, modes) must pass."""
    os.environ["PROBE_ONLY"] = "__none__"
    sys.argv = [sys.argv[0], "--quick"]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gpu_probe as gp
    for fn in (gp._gemm_shapes, gp._kfill, gp._chol, gp._lml, gp._sparse):
        gp.FAILS.clear()
        fn()
        assert not gp.FAILS, (fn.__name__, gp.FAILS)


def test_dense_kernels_entrywise(fv, golden):
    from fvgp_b200 import kernels as K
    from fvgp_b200.gp_prior import GPprior
    g = golden("dense_kernels")
    x1, x2, h = g["x1"], g["x2"], g["hps"]
    assert rel(np.asarray(GPprior._default_kernel(x1, x2, h)), g["default_12"]) <= 1e-12
    assert rel(np.asarray(GPprior._default_kernel(x1, x1, h)), g["default_11"]) <= 1e-12
    assert rel(np.asarray(K.get_distance_matrix(x1, x2)), g["d_iso"]) <= 1e-13
    assert rel(np.asarray(K.get_anisotropic_distance_matrix(x1, x2, h[1:])), g["d_ani"]) <= 1e-13
    for nm, f in (("se", K.squared_exponential_kernel), ("exp", K.exponential_kernel),
                  ("matern32", K.matern_kernel_diff1), ("matern52", K.matern_kernel_diff2)):
        assert rel(np.asarray(f(K.get_distance_matrix(x1, x2), 0.37)), g[nm + "_iso"]) <= 1e-12, nm
        assert rel(np.asarray(f(K.get_anisotropic_distance_matrix(x1, x2, h[1:]), 1.3)), g[nm + "_ani"]) <= 1e-12, nm
        assert rel(f(g["d_iso"], 0.37), g[nm + "_iso"]) <= 1e-12, nm          # plain-ndarray call path
    assert rel(np.asarray(GPprior._default_kernel(g["c1_x"], g["c1_x"], g["c1_hps"])), g["c1_K"]) <= 1e-12


def test_default_kernel_gradient_dense(fv, golden):
    from fvgp_b200 import GP
    g = golden("dense_kernels")
    x1, x2, h = g["x1"], g["x2"], g["hps"]
    gp = GP(x1, np.zeros(len(x1)) + np.arange(len(x1)) * 0.1, init_hyperparameters=h, noise_variances=np.full(len(x1), .1))
    d11 = gp.prior.dk_dh(x1, x1, h)
    d12 = gp.prior.dk_dh(x1, x2, h)
    assert np.max(np.abs(d11 - g["default_grad_11"])) <= 1e-12
    assert np.max(np.abs(d12 - g["default_grad_12"])) <= 1e-12
    assert np.max(np.abs(gp.prior.dk_dh(x1, x2, h, direction=2) - g["default_grad_12"][2])) <= 1e-12


@pytest.mark.parametrize("tag", ["c1", "c2"])
def test_dense_gp_against_reference(fv, golden, tag):
    from fvgp_b200 import GP
    g = golden("dense_lml_" + tag)
    x, y, nz = g["x"], g["y"], g["noise"]
    gp = GP(x, y, init_hyperparameters=g["h0"], noise_variances=nz)
    assert abs(gp.log_likelihood() / g["lml_h0_state"] - 1) <= 1e-8
    for hk in ("h0", "h1"):
        assert abs(gp.log_likelihood(g[hk]) / g["lml_" + hk] - 1) <= 1e-8
        assert rel(gp.neg_log_likelihood_gradient(g[hk]), g["grad_" + hk]) <= 1e-8
        assert abs(gp.neg_log_likelihood(g[hk]) + g["lml_" + hk]) <= 1e-8 * abs(g["lml_" + hk])
    assert rel(gp.kv.KVinvY, g["KVinvY_h0"]) <= 1e-7
    assert abs(gp.kv.logdet_KV / g["logdet_h0"] - 1) <= 1e-10
    assert rel(gp.K, fv.gp_prior.GPprior._default_kernel(x, x, g["h0"]).to_host()) == 0.0
    pm = gp.posterior_mean(g["x_pred"])
    assert np.allclose(pm["m(x)"], g["post_mean"], rtol=1e-8, atol=1e-10)
    pc = gp.posterior_covariance(g["x_pred"])
    assert np.allclose(pc["S"], g["post_S"], rtol=1e-6, atol=1e-9)
    assert np.allclose(pc["v(x)"], g["post_var"], rtol=1e-6, atol=1e-9)
    assert gp.posterior_covariance(g["x_pred"], variance_only=True)["S"] is None
    # analytic gradient agrees with finite differences of our own LML (gp_marginal_likelihood.py:338-364)
    fd, an = gp.test_log_likelihood_gradient(g["h1"])
    assert np.allclose(fd, an, rtol=2e-3, atol=1e-3)
    # default noise (gp_likelihood.py:102-104)
    gpd = GP(x[:200], y[:200], init_hyperparameters=g["h0"])
    assert abs(gpd.log_likelihood(g["h1"]) / g["lml_default_noise"] - 1) <= 1e-8
    # user kernel composed from fvgp.kernels names (examples' skernel): fused lazily, same LML
    from fvgp_b200.kernels import get_distance_matrix, squared_exponential_kernel

    def se(x1, x2, h):
        return h[0] * squared_exponential_kernel(get_distance_matrix(x1, x2), h[1])
    gps = GP(x[:300], y[:300], init_hyperparameters=g["se_hps"], noise_variances=nz[:300], kernel_function=se)
    assert abs(gps.log_likelihood(g["se_hps"] * 1.1) / g["lml_se"] - 1) <= 1e-8
    # ... and its gradient through finite-difference dK (gp_prior.py:438-447) stays consistent with FD of the LML
    gr = gps.neg_log_likelihood_gradient(g["se_hps"])
    fd, _ = gps.test_log_likelihood_gradient(g["se_hps"])
    assert np.allclose(-gr, fd, rtol=5e-2, atol=5e-2)


def test_set_hyperparameters_update_data_and_pickle(fv, golden):
    from fvgp_b200 import GP
    g = golden("dense_lml_c2")
    x, y, nz = g["x"], g["y"], g["noise"]
    gp = GP(x[:400], y[:400], init_hyperparameters=g["h0"], noise_variances=nz[:400])
    gp.update_gp_data(x[400:437], y[400:437], noise_variances_new=nz[400:437], append=True)   # bordered Cholesky update
    assert gp.kv.state.info.get("appended_rows") == 37
    gp.update_gp_data(x[437:], y[437:], noise_variances_new=nz[437:], append=True)
    assert gp.kv.state.info.get("appended_rows") == len(x) - 437
    assert abs(gp.log_likelihood() / g["lml_h0"] - 1) <= 1e-8
    fresh = GP(x, y, init_hyperparameters=g["h0"], noise_variances=nz)
    assert abs(gp.kv.logdet_KV / fresh.kv.logdet_KV - 1) <= 1e-12
    assert rel(gp.kv.KVinvY, fresh.kv.KVinvY) <= 1e-7
    assert np.max(np.abs(np.tril(gp.kv.Chol_factor) - np.tril(fresh.kv.Chol_factor))) <= 1e-11
    assert np.allclose(gp.posterior_covariance(g["x_pred"])["v(x)"], g["post_var"], rtol=1e-6, atol=1e-9)
    gp.set_hyperparameters(g["h1"])
    assert abs(gp.log_likelihood() / g["lml_h1"] - 1) <= 1e-8
    K0, m0 = gp.K.copy(), gp.posterior_mean(g["x_pred"])["m(x)"]
    gp2 = pickle.loads(pickle.dumps(gp))                                   # tests/test_fvgp.py:1108
    assert np.array_equal(gp2.K, K0) and np.array_equal(gp2.V, gp.V)
    assert np.allclose(gp2.posterior_mean(g["x_pred"])["m(x)"], m0, rtol=1e-12)


def test_linalg_modes_and_custom_callables(fv, golden):
    """tests/test_fvgp.py:357, :4188, :5030."""
    import scipy.linalg as sla
    from fvgp_b200 import GP
    g = golden("dense_lml_c1")
    x, y, nz = g["x"][:200], g["y"][:200], g["noise"][:200]
    base = GP(x, y, init_hyperparameters=g["h0"], noise_variances=nz)
    ref = base.log_likelihood(g["h1"])
    for mode in ("CholInv", "Inv"):
        gp = GP(x, y, init_hyperparameters=g["h0"], noise_variances=nz, linalg_mode=mode)
        assert abs(gp.log_likelihood(g["h1"]) / ref - 1) <= 1e-10
        assert np.allclose(gp.kv.KVinv @ gp.kv.KV, np.eye(200), atol=1e-8)
    calls = (lambda KV: sla.cho_factor(KV, lower=True), lambda f, b: sla.cho_solve(f, b),
             lambda f: 2 * np.sum(np.log(np.diag(f[0]))))
    gp = GP(x, y, init_hyperparameters=g["h0"], noise_variances=nz, linalg_mode=calls)
    assert abs(gp.log_likelihood(g["h1"]) / ref - 1) <= 1e-10
    assert rel(gp.neg_log_likelihood_gradient(g["h1"]), base.neg_log_likelihood_gradient(g["h1"])) <= 1e-7
    with pytest.raises(Exception):
        GP(x, y, init_hyperparameters=g["h0"], noise_variances=nz, linalg_mode="nonsense")


def test_non_positive_definite_raises(fv):
    from fvgp_b200 import GP

    def bad_kernel(x1, x2, h):
        return -np.ones((len(x1), len(x2)))
    with pytest.raises(Exception, match="positive definite"):
        GP(np.random.rand(70, 2), np.random.rand(70), init_hyperparameters=np.ones(3), noise_variances=np.full(70, 1e-3),
           kernel_function=bad_kernel)


def test_multitask_against_reference(fv, golden):
    from fvgp_b200 import fvGP
    g = golden("multitask")
    gp = fvGP(g["x"], g["y"], init_hyperparameters=g["h"], noise_variances=g["noise"])
    assert np.array_equal(gp.x_data, g["x_index"]) and np.array_equal(gp.y_data, g["y_flat"])
    assert abs(gp.log_likelihood(g["h1"]) / g["lml"] - 1) <= 1e-8
    assert rel(gp.neg_log_likelihood_gradient(g["h1"]), g["grad"]) <= 1e-8
    pm = gp.posterior_mean(g["x_pred"])
    assert np.allclose(pm["m(x)"], g["post_mean"], rtol=1e-8, atol=1e-10)
    assert np.allclose(pm["m(x)_flat"], g["post_mean_flat"], rtol=1e-8, atol=1e-10)
    pc = gp.posterior_covariance(g["x_pred"])
    assert pc["v(x)"].shape == (7, 3) and pc["S"].shape == (7, 7, 3, 3)


@pytest.mark.parametrize("tag", ["t3152", "c4small"])
def test_gp2scale_against_reference(fv, golden, tag):
    from fvgp_b200 import GP
    g = golden("gp2scale_" + tag)
    x, y, nz = g["x"], g["y"], g["noise"]
    gp = GP(x, y, init_hyperparameters=g["h0"], noise_variances=nz, gp2Scale=True, linalg_mode="sparseLU")
    K = gp.K
    assert K.indices.dtype == np.int32 and K.has_sorted_indices
    assert np.array_equal(K.indptr, g["indptr"]) and np.array_equal(K.indices, g["indices"])      # bit-exact
    if "data" in g:
        assert rel(K.data, g["data"]) <= 1e-12
    else:
        assert rel(K.data[:2000], g["data_head"]) <= 1e-12 and abs(K.data.sum() / g["data_sum"] - 1) <= 1e-13
    K1 = gp.prior.compute_prior_covariance_matrix(x, g["h1"])
    assert np.array_equal(K1.indptr, g["h1_indptr"]) and np.array_equal(K1.indices, g["h1_indices"])
    assert (K != K.T).nnz == 0
    for hk in ("h0", "h1"):
        assert abs(gp.log_likelihood(g[hk]) / g["lml_" + hk] - 1) <= 1e-8
    assert rel(gp.kv.KVinvY, g["KVinvY_h0"]) <= 1e-6
    assert np.allclose(gp.posterior_mean(g["x_pred"])["m(x)"], g["post_mean"], rtol=1e-7, atol=1e-9)
    assert np.allclose(gp.posterior_covariance(g["x_pred"])["v(x)"], g["post_var"], rtol=1e-6, atol=1e-8)
    with pytest.raises(Exception, match="gp2Scale"):
        gp.neg_log_likelihood_gradient(g["h0"])                       # gp_marginal_likelihood.py:240
    # Krylov modes: tight CG reproduces the exact solve; SLQ logdet within its own error bars
    for mode in ("sparseCG", "sparseCGpre"):
        gpk = GP(x, y, init_hyperparameters=g["h0"], noise_variances=nz, gp2Scale=True, linalg_mode=mode,
                 args={"sparse_cg_tol": 1e-11, "random_logdet_lanczos_degree": 30,
                       "random_logdet_min_num_samples": 40, "random_logdet_error_rtol": 0.003})
        assert np.allclose(gpk.kv.KVinvY, g["KVinvY_h0"], rtol=1e-6, atol=1e-8 * np.abs(g["KVinvY_h0"]).max())
        sd = np.sqrt(gpk.kv.last_logdet_variance)
        assert abs(gpk.kv.logdet_KV - g["logdet_h0"]) <= 5 * sd + 0.01 * abs(g["logdet_h0"])
        assert gpk.marginal_likelihood.log_likelihood_variance() is not None


def test_gp2scale_edge_cases(fv):
    from fvgp_b200 import _lib as L, ops
    h = np.array([2.0, .4, .35])
    far = ops.wendland_csr(L.to_dev(np.zeros((5, 2))), L.to_dev(np.ones((4, 2)) * 10), h)
    assert far.nnz == 0 and far.to_scipy().shape == (5, 4)                 # disjoint boxes (test_fvgp.py:1737)
    one = L.to_dev(np.array([[0.3, 0.3]]))
    K = ops.wendland_csr(one, one, h).to_scipy()
    assert K.nnz == 1 and K[0, 0] == 2.0                                   # diagonal == amplitude (test_fvgp.py:1724)
    dup = L.to_dev(np.tile(np.array([[0.5, 0.5]]), (70, 1)))               # coincident points: dense 70x70 block
    K = ops.wendland_csr(dup, dup, h).to_scipy()
    assert K.nnz == 4900 and np.all(K.data == 2.0)


def test_size_independent_properties_at_scale(fv):
    """Properties that need no oracle: symmetry, L L^T = KV, KV KV^-1 = I, mirrored sparsity."""
    import torch
    from fvgp_b200 import _lib as L, ops
    rng = np.random.default_rng(9)
    n = 6000
    x = L.to_dev(rng.random((n, 3)))
    noise = L.to_dev(np.full(n, 1e-2))
    h = np.array([1.0, .3, .4, .5])
    full, _ = ops.kfill(L.K_MATERN32, x, x, h[0], 1 / h[1:], 1.0, noise=noise, mode=L.FILL_SYMMETRIC)
    KV = full[:, :n].clone()
    assert torch.equal(KV, KV.T)                                            # mirror tiles are bitwise transposes
    ref, _ = ops.kfill(L.K_MATERN32, x, x, h[0], 1 / h[1:], 1.0, noise=noise, mode=L.FILL_FULL)
    assert torch.equal(KV, ref[:, :n])
    f = ops.potrf(full, _, n)
    Lw = f.lower()
    assert float((Lw @ Lw.T - KV).abs().max()) <= 1e-12 * float(KV.abs().max()) * 50
    ops.potri(f)
    low = torch.tril(f.buf[:, :n])
    Kinv = low + torch.tril(low, -1).T
    assert float((Kinv @ KV - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max()) <= 1e-9
    ns = 50000
    xs = rng.random((ns, 3))
    xs = L.to_dev(xs[np.lexsort((xs[:, 2] // .1, xs[:, 1] // .1, xs[:, 0] // .1))])
    A = ops.wendland_csr(xs, xs, np.array([1.0, .05, .05, .05])).to_scipy()
    assert (A != A.T).nnz == 0 and np.all(A.diagonal() == 1.0) and A.has_sorted_indices
    shuffled = ops.wendland_csr(xs.flip(0).contiguous(), xs.flip(0).contiguous(), np.array([1.0, .05, .05, .05])).to_scipy()
    assert shuffled.nnz == A.nnz                                            # ordering changes the layout, not the set


def _random_spd_csr(rng, n, row_nnz):
    """Symmetric, diagonally dominant CSR with ragged rows (some empty off-diagonals, some > 128 entries)."""
    import scipy.sparse as sp
    rows, cols, vals = [], [], []
    for i in range(n):
        k = int(row_nnz[i])
        c = rng.choice(n, size=min(k, n), replace=False)
        rows += [i] * len(c)
        cols += list(c)
        vals += list(rng.standard_normal(len(c)))
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    A = A + A.T
    A = A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)
    A = A.tocsr()
    A.sort_indices()
    return A


def _to_device_csr(A):
    import torch
    from fvgp_b200 import ops
    return ops.DeviceCSR(torch.as_tensor(A.indptr.astype(np.int64)).cuda(), torch.as_tensor(A.indices.astype(np.int32)).cuda(),
                         torch.as_tensor(A.data).cuda(), A.shape)


def test_spmv_pcg_slq_on_ragged_matrix(fv):
    """SpMV (4 strips in flight), the 3-kernel PCG and the batched Lanczos against scipy on a ragged SPD matrix."""
    from fvgp_b200 import _lib as L, ops
    from oracle import fvgp_oracle as orc
    rng = np.random.default_rng(21)
    n = 3001
    row_nnz = rng.integers(0, 40, n)
    row_nnz[::97] = 300                                   # rows longer than one 128-entry sweep
    row_nnz[5:9] = 0
    A = _random_spd_csr(rng, n, row_nnz)
    Ad = _to_device_csr(A)
    v = rng.standard_normal(n)
    y = ops.spmv(Ad, L.to_dev(v)).cpu().numpy()
    assert np.max(np.abs(y - A @ v)) <= 1e-13 * np.max(np.abs(A @ v))
    b = rng.standard_normal(n)
    for pre in (None, ops.bjacobi(Ad)):
        for rtol in (1e-5, 1e-11):
            x, info, iters, relres = ops.pcg(Ad, L.to_dev(b), rtol=rtol, precond=pre)
            assert info == 0 and relres < rtol
            assert np.linalg.norm(A @ x.cpu().numpy() - b) < rtol * np.linalg.norm(b) * 1.0001
    ref, ref_iters = orc.sparse_cg(A, b[:, None], rtol=1e-5)
    x, info, iters, relres = ops.pcg(Ad, L.to_dev(b), rtol=1e-5)
    assert abs(iters - ref_iters[0]) <= 1                  # same recurrence and stopping rule as scipy cg
    assert rel(x.cpu().numpy(), ref[:, 0]) <= 1e-4
    _, info, iters, _ = ops.pcg(Ad, L.to_dev(b), rtol=1e-14, maxiter=3)
    assert info == 1 and iters == 3                        # scipy: info > 0 when maxiter is reached
    x0 = np.linalg.solve(A.toarray(), b)
    _, info, iters, _ = ops.pcg(Ad, L.to_dev(b), x0=L.to_dev(x0), rtol=1e-8)
    assert info == 0 and iters == 0                        # converged warm start: no update applied
    # batched Lanczos (16 + 8 + 2 + 1 probes) == the same probes one at a time
    est, var, samples = ops.slq_logdet(Ad, degree=25, probes=27, seed=3)
    singles = np.array([ops.slq_logdet(Ad, degree=25, probes=1, seed=3, probe0=p)[2][0] for p in (0, 7, 15, 16, 23, 24, 26)])
    assert rel(samples[[0, 7, 15, 16, 23, 24, 26]], singles) <= 1e-9
    exact = np.linalg.slogdet(A.toarray())[1]
    assert abs(est - exact) <= 6 * np.sqrt(var) + 0.01 * abs(exact)


def test_gp2scale_value_underflow_pattern(fv):
    """np.nonzero drops entries whose VALUE rounds to zero (gp2Scale_covariance.py:147): with a tiny
    amplitude the near-edge pairs disappear from the pattern; the strict kernel variant replays that."""
    from fvgp_b200 import _lib as L, ops
    from oracle import fvgp_oracle as orc
    rng = np.random.default_rng(8)
    x = np.concatenate([rng.random((300, 1)), np.array([[0.0], [1.0 - 2.0 ** -20], [1.0 - 2.0 ** -30]])])
    for amp in (1e-300, 1e-250, 1e-120, 3.0):
        h = np.array([amp, 1.0])
        ref = orc.gp2scale_covariance(x, x, h, batch=100, symmetric=True)
        K = ops.wendland_csr(L.to_dev(x), L.to_dev(x), h).to_scipy()
        assert np.array_equal(K.indptr, ref.indptr) and np.array_equal(K.indices, ref.indices), amp
        normal = np.abs(ref.data) > 1e-290                            # subnormal values carry fewer digits
        assert rel(K.data[normal], ref.data[normal]) <= 1e-12
    assert ref.nnz > orc.gp2scale_covariance(x, x, np.array([1e-300, 1.0]), batch=100, symmetric=True).nnz


def test_gp2scale_sliver_pairs(fv):
    """Pairs whose s is within rounding of 1: the approximate test may not decide them, the exact sequence does."""
    from fvgp_b200 import _lib as L, ops
    from oracle import fvgp_oracle as orc
    th = np.array([1.0, 0.3, 0.7])
    base = np.array([0.25, 0.5])
    ang = np.linspace(0, 2 * np.pi, 400, endpoint=False)
    ring = base + np.stack([th[1] * np.cos(ang), th[2] * np.sin(ang)], axis=1)      # s == 1 up to rounding
    pts = [base[None, :], ring]
    for k in range(1, 4):
        pts.append(np.nextafter(ring, base[None, :] + 0 * ring) if k == 1 else ring * (1 + (k - 2) * 2.0 ** -52))
    x = np.concatenate(pts)
    ref = orc.gp2scale_covariance(x, x, th, batch=500, symmetric=True)
    K = ops.wendland_csr(L.to_dev(x), L.to_dev(x), th).to_scipy()
    assert np.array_equal(K.indptr, ref.indptr) and np.array_equal(K.indices, ref.indices)
    assert 0 < ref[0].nnz < len(x)                               # the ring really straddles the support edge


def test_potrf_lookahead_ragged_size(fv):
    """n >= 4 * 2048 takes the two-stream look-ahead factorisation; ragged last block; vs cuSOLVER."""
    import torch
    from fvgp_b200 import _lib as L, ops
    rng = np.random.default_rng(12)
    n = 8192 + 1111
    x = L.to_dev(rng.random((n, 3)))
    noise = L.to_dev(np.full(n, 1e-2))
    h = np.array([1.0, .3, .4, .5])
    buf, ld = ops.kfill(L.K_MATERN32, x, x, h[0], 1 / h[1:], 1.0, noise=noise, mode=L.FILL_SYMMETRIC)
    ref = torch.linalg.cholesky(buf[:, :n])
    f = ops.potrf(buf, ld, n)
    err = float((f.lower() - ref).abs().max() / ref.abs().max())
    assert err <= 1e-11, err
    rhs = L.to_dev(rng.standard_normal((1, n)))
    sol = ops.potrs(f, rhs.clone())
    ref_sol = torch.cholesky_solve(rhs.T.contiguous(), ref).T
    assert float((sol - ref_sol).abs().max() / ref_sol.abs().max()) <= 1e-8
    # failure inside a look-ahead panel is still reported with its global pivot index
    bad, ld = ops.kfill(L.K_MATERN32, x, x, h[0], 1 / h[1:], 1.0, noise=noise, mode=L.FILL_SYMMETRIC)
    bad[5000, 5000] = -1.0
    with pytest.raises(L.NonPositiveDefiniteError) as ei:
        ops.potrf(bad, ld, n)
    assert ei.value.pivot == 5001


def test_fused_gradient_for_named_user_kernels(fv):
    """User kernels composed of the fvgp.kernels names get the fused gradient traces (no dK materialised); checked
    against grad(-LML)_h = -1/2 (b^T dK_h b - tr(KV^-1 dK_h)) formed on the host with numpy and finite-difference dK."""
    from scipy.spatial.distance import cdist
    from fvgp_b200 import GP
    from fvgp_b200.kernels import (exponential_kernel, get_anisotropic_distance_matrix, get_distance_matrix,
                                   matern_kernel_diff1, matern_kernel_diff2, squared_exponential_kernel)
    rng = np.random.default_rng(17)
    n = 400
    x = rng.random((n, 2))
    y = np.sin(4 * x[:, 0]) + np.cos(3 * x[:, 1]) + 0.05 * rng.standard_normal(n)
    nz = np.full(n, 2e-2)

    def np_kernel(name, x1, x2, h):
        if name == "se_iso":
            return h[0] * np.exp(-cdist(x1, x2) ** 2 / (2 * h[1] ** 2))
        if name == "m52_aniso":
            d = cdist(x1 / h[1:3], x2 / h[1:3])
            return h[0] ** 2 * (1 + np.sqrt(5) * d + 5 * d ** 2 / 3) * np.exp(-np.sqrt(5) * d)
        if name == "exp_iso":
            return h[0] * np.exp(-cdist(x1, x2) / h[1])
        d = cdist(x1 / h[1:3], x2 / h[1:3])                                   # m32 with a length parameter on top
        return h[0] * (1 + np.sqrt(3) * d / h[3]) * np.exp(-np.sqrt(3) * d / h[3])

    kernels = {
        "se_iso": (lambda a, b, h: h[0] * squared_exponential_kernel(get_distance_matrix(a, b), h[1]), np.array([1.3, 0.4])),
        "m52_aniso": (lambda a, b, h: h[0] ** 2 * matern_kernel_diff2(get_anisotropic_distance_matrix(a, b, h[1:3]), 1.0),
                      np.array([1.1, 0.35, 0.6])),
        "exp_iso": (lambda a, b, h: h[0] * exponential_kernel(get_distance_matrix(a, b), h[1]), np.array([0.9, 0.7])),
        "m32_len": (lambda a, b, h: h[0] * matern_kernel_diff1(get_anisotropic_distance_matrix(a, b, h[1:3]), h[3]),
                    np.array([1.2, 0.5, 0.8, 0.9])),
    }
    for name, (kern, h) in kernels.items():
        gp = GP(x, y, init_hyperparameters=h, noise_variances=nz, kernel_function=kern)
        assert gp.marginal_likelihood._descriptor(h) is not None, name       # the fused path is taken
        grad = gp.neg_log_likelihood_gradient(h)
        KV = np_kernel(name, x, x, h) + np.diag(nz)
        assert abs(gp.log_likelihood(h) / (-0.5 * ((y - y.mean()) @ np.linalg.solve(KV, y - y.mean())
                                                    + np.linalg.slogdet(KV)[1] + n * np.log(2 * np.pi))) - 1) <= 1e-9, name
        b = np.linalg.solve(KV, y - y.mean())
        Kinv = np.linalg.inv(KV)
        ref = np.zeros(len(h))
        for i in range(len(h)):
            hp, hm = h.copy(), h.copy()
            hp[i] += 1e-6
            hm[i] -= 1e-6
            dK = (np_kernel(name, x, x, hp) - np_kernel(name, x, x, hm)) / 2e-6
            ref[i] = -0.5 * (b @ dK @ b - np.sum(Kinv * dK))
        assert np.max(np.abs(grad - ref) / np.abs(ref)) <= 2e-6, (name, grad, ref)


def _pop_problem(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    y = np.sin(5 * x[:, 0]) + (np.cos(3 * x[:, 1]) if d > 1 else 0.0) + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2)


def test_population_matches_one_at_a_time(fv):
    """B proposals on concurrent streams (fvgp_lml_population) give, proposal by proposal, the LML of
    GP.log_likelihood to the last bit (same kernels, same order, quadratic form summed on the host the same way) and
    the gradient of GP.neg_log_likelihood_gradient; both are also pinned to the oracle (1e-8)."""
    from fvgp_b200 import GP
    from oracle import fvgp_oracle as orc
    # n >= 6144 takes the stream schedule by itself (each proposal runs the look-ahead POTRF on its own stream)
    for n, d, B in ((700, 3, 9), (1000, 1, 40), (2500, 2, 5), (6200, 2, 3)):
        x, y, nz = _pop_problem(n, d, 100 + n)
        h0 = np.array([1.0] + [0.3] * d)
        gp = GP(x, y, init_hyperparameters=h0, noise_variances=nz)
        assert gp.marginal_likelihood.population_supported(want_grad=True)
        rng = np.random.default_rng(n)
        T = h0 * (0.7 + 0.6 * rng.random((B, d + 1)))
        lml_pop, grad_pop = gp.marginal_likelihood.evaluate_population(T, with_gradient=True)
        lml_only = gp.log_likelihood_population(T)
        os.environ["FVGP_POPULATION_STREAMS"] = "1"               # the stream schedule (what n >= 6144 uses)
        try:
            lml_st, grad_st = gp.marginal_likelihood.evaluate_population(T, with_gradient=True)
        finally:
            os.environ.pop("FVGP_POPULATION_STREAMS")
        assert np.array_equal(lml_st, lml_pop) and rel(grad_st, grad_pop) <= 1e-12
        for b in range(B):
            one = gp.log_likelihood(T[b])
            assert lml_pop[b] == one and lml_only[b] == one, (n, b, lml_pop[b], lml_only[b], one)
            g = gp.neg_log_likelihood_gradient(T[b])
            assert rel(grad_pop[b], g) <= 1e-9, (n, b, grad_pop[b], g)
        for b in (0, B - 1):
            assert abs(lml_pop[b] / orc.dense_log_likelihood(x, y, T[b], nz) - 1) <= 1e-8
            assert rel(grad_pop[b], orc.dense_neg_log_likelihood_gradient(x, y, T[b], nz, economical=True)) <= 1e-8


def test_population_user_kernel_hessian_and_failures(fv):
    from fvgp_b200 import GP
    from fvgp_b200.kernels import get_anisotropic_distance_matrix, matern_kernel_diff2
    x, y, nz = _pop_problem(600, 2, 7)
    kern = lambda a, b, h: h[0] * matern_kernel_diff2(get_anisotropic_distance_matrix(a, b, h[1:3]), 1.0)   # noqa: E731
    h = np.array([1.2, 0.4, 0.5])
    gp = GP(x, y, init_hyperparameters=h, noise_variances=nz, kernel_function=kern)
    ml = gp.marginal_likelihood
    assert ml.population_supported(want_grad=True)
    T = h * np.array([[1.0, 1.0, 1.0], [1.1, 0.9, 1.2], [0.8, 1.3, 0.7]])
    lml, grad = ml.evaluate_population(T, with_gradient=True)
    for b in range(3):
        assert lml[b] == gp.log_likelihood(T[b])
        assert rel(grad[b], gp.neg_log_likelihood_gradient(T[b])) <= 1e-8
    # Hessian: H + 1 gradients as one population == the reference's forward differences of single gradients
    Hm = gp.neg_log_likelihood_hessian(h)
    g0 = gp.neg_log_likelihood_gradient(h)
    for i in range(3):
        hp = h.copy()
        hp[i] += 1e-6
        row = (gp.neg_log_likelihood_gradient(hp) - g0) / 1e-6
        assert np.allclose(Hm[i, i:], row[i:], rtol=1e-4, atol=1e-4 * np.abs(Hm).max())
    assert np.array_equal(Hm, Hm.T)
    # a proposal whose matrix is not positive definite is reported like the one-at-a-time path (raises)
    bad = np.array([[1.0, 0.4, 0.5], [-1.0, 0.4, 0.5]])                  # negative amplitude: first pivot < 0
    with pytest.raises(Exception, match="Linear algebra failed"):
        gp.log_likelihood_population(bad)
    with pytest.raises(Exception, match="Linear algebra failed"):
        gp.log_likelihood(bad[1])
    assert lml[0] == gp.log_likelihood_population(T)[0]                   # the evaluator is usable afterwards
    # kernels that do not fold into a fused radial expression fall back to one-at-a-time evaluation
    gpa = GP(x, y, init_hyperparameters=h, noise_variances=nz,
             kernel_function=lambda a, b, hh: np.asarray(kern(a, b, hh)) + 0.0)
    assert not gpa.marginal_likelihood.population_supported()
    assert rel(gpa.log_likelihood_population(T), lml) <= 1e-12


def test_train_global_uses_the_population_path(fv):
    from fvgp_b200 import GP, ops
    x, y, nz = _pop_problem(500, 1, 3)
    gp = GP(x, y, init_hyperparameters=np.array([1.0, 0.3]), noise_variances=nz)
    ops.start_phase_timing()
    before = gp.log_likelihood()
    hps = gp.train(hyperparameter_bounds=np.array([[0.05, 10.0], [0.02, 5.0]]), method="global", max_iter=6, pop_size=8)
    phases = ops.stop_phase_timing()
    assert "population" in phases and gp.log_likelihood() >= before
    assert np.all(hps >= [0.05, 0.02]) and np.all(hps <= [10.0, 5.0])


def test_robust_and_wendland_kernel_names(fv, golden):
    """the *_robust radial kernels, wendland_kernel, the dense wendland_anisotropic and
    the support-aware sparse block kernel, against outputs of the unmodified reference (tests/golden/robust_kernels.npz)."""
    from fvgp_b200 import kernels as K
    g = golden("robust_kernels")
    x1, x2, h, phi = g["x1"], g["x2"], g["hps"], float(g["phi"])
    d_iso = np.asarray(K.get_distance_matrix(x1, x2))
    d_ani = np.asarray(K.get_anisotropic_distance_matrix(x1, x2, h[1:]))
    for nm, f in (("se", K.squared_exponential_kernel_robust), ("exp", K.exponential_kernel_robust),
                  ("matern32", K.matern_kernel_diff1_robust), ("matern52", K.matern_kernel_diff2_robust)):
        assert rel(np.asarray(f(K.get_distance_matrix(x1, x2), phi)), g[nm + "_robust_iso"]) <= 1e-12, nm      # fused
        assert rel(np.asarray(f(K.get_anisotropic_distance_matrix(x1, x2, h[1:]), phi)), g[nm + "_robust_ani"]) <= 1e-12, nm
        assert rel(f(d_iso, phi), g[nm + "_robust_iso"]) <= 1e-12, nm                                          # ndarray
        assert rel(1.7 * np.asarray(f(K.get_distance_matrix(x1, x2), 0.0)), np.full(d_iso.shape, 1.7)) <= 1e-15
    dd = np.array(d_ani, copy=True)
    w = K.wendland_kernel(dd)
    assert np.max(np.abs(w - g["wendland_kernel_ani"])) <= 1e-13 and dd.max() <= 1.0       # clamped in place, like the reference
    assert np.max(np.abs(np.asarray(K.wendland_anisotropic(x1, x2, h)) - g["wendland_anisotropic_12"])) <= 1e-13
    assert np.max(np.abs(np.asarray(K.wendland_anisotropic(x1, x1, h)) - g["wendland_anisotropic_11"])) <= 1e-13
    for f in (K.wendland_anisotropic_gp2Scale_cpu, K.wendland_anisotropic_gp2Scale_cpu_sparse,
              K.wendland_anisotropic_gp2Scale_gpu_sparse):
        blk = f(x1, x2, h).toarray()
        assert np.array_equal(blk != 0, g["wendland_block_12"] != 0)                     # pattern bit-exact
        assert rel(blk[blk != 0], g["wendland_block_12"][blk != 0]) <= 1e-12              # values (K entries: 1e-12)
        assert np.max(np.abs(blk - g["wendland_sparse_12"])) <= 1e-12                     # KD-tree variant (a17)


def test_fused_substitution_and_two_column_targets(fv):
    """The single-launch substitution kernel (n <= 2048) against the step-per-launch kernels it replaces (bitwise: same
    arithmetic per row / column), for ragged sizes; and y_data with two columns (r = 2 right-hand sides, the quadratic
    form averaged over columns, gp_marginal_likelihood.py:171-178) through the one-at-a-time and the population path."""
    from fvgp_b200 import GP
    from oracle import fvgp_oracle as orc
    for n in (37, 64, 129, 700, 2048):
        x, y, nz = _pop_problem(n, 2, 900 + n)
        h = np.array([1.1, 0.35, 0.45])
        gp = GP(x, y, init_hyperparameters=h, noise_variances=nz)
        fused = gp.log_likelihood(h * 1.01), gp.kv.evaluate(h * 1.01, gp.likelihood.V, gp.prior.m).KVinvY.copy()
        os.environ["FVGP_TRSV_FUSED"] = "0"
        try:
            gp.kv._memo = None
            steps = gp.log_likelihood(h * 1.01), gp.kv.evaluate(h * 1.01, gp.likelihood.V, gp.prior.m).KVinvY.copy()
        finally:
            os.environ.pop("FVGP_TRSV_FUSED")
            gp.kv._memo = None
        assert fused[0] == steps[0] and np.array_equal(fused[1], steps[1]), n
        assert abs(fused[0] / orc.dense_log_likelihood(x, y, h * 1.01, nz) - 1) <= 1e-8
    x, y, nz = _pop_problem(900, 2, 5)
    y2 = np.column_stack([y, np.cos(4 * x[:, 0]) - 0.3 * x[:, 1]])
    h = np.array([1.0, 0.3, 0.4])
    gp = GP(x, y2, init_hyperparameters=h, noise_variances=nz)
    T = h * np.array([[1.0, 1.0, 1.0], [1.2, 0.9, 1.1], [0.9, 1.1, 0.8]])
    lml, grad = gp.marginal_likelihood.evaluate_population(T, with_gradient=True, component=1)
    for b in range(3):
        assert lml[b] == gp.log_likelihood(T[b])
        assert abs(lml[b] / orc.dense_log_likelihood(x, y2, T[b], nz) - 1) <= 1e-8
        assert rel(grad[b], gp.neg_log_likelihood_gradient(T[b], component=1)) <= 1e-9
        assert rel(grad[b], orc.dense_neg_log_likelihood_gradient(x, y2, T[b], nz, component=1, economical=True)) <= 1e-8


def test_train_mcmc_speculates_through_the_population_path(fv):
    """train(method="mcmc"), the reference's default: batches of proposals around the current state go through
    fvgp_lml_population (gp_training.run_mcmc); the chain must still find the posterior mode of the hyperparameters."""
    from fvgp_b200 import GP, ops
    x, y, nz = _pop_problem(600, 1, 11)
    gp = GP(x, y, init_hyperparameters=np.array([0.4, 0.9]), noise_variances=nz)
    before = gp.log_likelihood()
    ops.start_phase_timing()
    hps = gp.train(hyperparameter_bounds=np.array([[0.05, 10.0], [0.02, 5.0]]), method="mcmc", max_iter=150,
                   mcmc_args={"seed": 3})
    phases = ops.stop_phase_timing()
    info = gp.trainer.mcmc_info
    assert "population" in phases and info["likelihood calls"] < 150 and len(info["x"]) == 151
    assert gp.log_likelihood() >= before and np.all(hps > 0)
    assert gp.log_likelihood(info["max x"]) == info["max f(x)"]          # population LMLs are the one-at-a-time LMLs
