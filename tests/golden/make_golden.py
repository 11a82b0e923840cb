"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py

Inputs are seeded; outputs are what fvgp (reference) itself returns.  The fixtures pin
the numpy oracle (tests/test_oracle_golden.py) and, through it and directly, the CUDA
path (tests/test_gpu_*.py).  Seeds/shapes follow the reference's own result-pinning
tests where they exist (tests/test_fvgp.py:1711, :1724, :3027, :3152,
:3112) and SURVEY section 8d's configs at parity size otherwise.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

warnings.filterwarnings("ignore")
fv = ref_shim.install()
from fvgp import kernels as rk  # noqa: E402
from fvgp.gp_prior import GPprior  # noqa: E402
from fvgp import gp2Scale_covariance as g2  # noqa: E402


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("wrote", name, {k: np.shape(v) for k, v in arrays.items()})


def dense_kernels():
    rng = np.random.default_rng(101)
    x1 = rng.random((37, 3)) * 2.0 - 0.5
    x2 = rng.random((29, 3)) * 2.0 - 0.5
    hps = np.array([1.7, 0.3, 0.45, 0.8])
    out = dict(x1=x1, x2=x2, hps=hps)
    out["default_12"] = GPprior._default_kernel(x1, x2, hps)
    out["default_11"] = GPprior._default_kernel(x1, x1, hps)
    out["default_grad_11"] = GPprior._default_kernel_analytical_gradient(x1, x1, hps)
    out["default_grad_12"] = GPprior._default_kernel_analytical_gradient(x1, x2, hps)
    d_iso = rk.get_distance_matrix(x1, x2)
    d_ani = rk.get_anisotropic_distance_matrix(x1, x2, hps[1:])
    out["d_iso"], out["d_ani"] = d_iso, d_ani
    for nm, f in (("se", rk.squared_exponential_kernel), ("exp", rk.exponential_kernel),
                  ("matern32", rk.matern_kernel_diff1), ("matern52", rk.matern_kernel_diff2)):
        out[nm + "_iso"] = f(d_iso, 0.37)
        out[nm + "_ani"] = f(d_ani, 1.3)
    # 1-D input, config C1 shape
    xa = np.random.default_rng(1).random((64, 1))
    out["c1_x"] = xa
    out["c1_hps"] = np.array([1.3, 0.25])
    out["c1_K"] = GPprior._default_kernel(xa, xa, out["c1_hps"])
    save("dense_kernels", **out)


def dense_lml():
    for tag, n, d, seed in (("c1", 500, 1, 1), ("c2", 600, 3, 2)):
        rng = np.random.default_rng(seed)
        x = rng.random((n, d))
        if d == 1:
            y = np.sin(5 * x[:, 0]) + np.cos(10 * x[:, 0]) + 0.05 * rng.standard_normal(n)
            h0, h1 = np.array([1.0, 0.3]), np.array([1.3, 0.25])
        else:
            y = np.sin(5 * x[:, 0]) * np.cos(3 * x[:, 1]) + x[:, 2] + 0.1 * rng.standard_normal(n)
            h0, h1 = np.array([1.0, .3, .4, .5]), np.array([1.0, .3, .4, .5]) * 1.04
        noise = np.full(n, 1e-2)
        gp = fv.GP(x, y, init_hyperparameters=h0, noise_variances=noise)
        out = dict(x=x, y=y, noise=noise, h0=h0, h1=h1)
        out["lml_h0_state"] = gp.log_likelihood()
        out["lml_h0"] = gp.log_likelihood(h0)
        out["lml_h1"] = gp.log_likelihood(h1)
        out["grad_h0"] = gp.neg_log_likelihood_gradient(h0)
        out["grad_h1"] = gp.neg_log_likelihood_gradient(h1)
        out["KVinvY_h0"] = gp.kv.KVinvY
        out["logdet_h0"] = gp.kv.logdet_KV
        xp = np.random.default_rng(seed + 50).random((11, d))
        pm = gp.posterior_mean(xp)
        pc = gp.posterior_covariance(xp)
        out["x_pred"], out["post_mean"], out["post_var"], out["post_S"] = xp, pm["m(x)"], pc["v(x)"], pc["S"]
        # default noise (no noise_variances): (mean|y|/100)^2
        gpd = fv.GP(x[:200], y[:200], init_hyperparameters=h0)
        out["lml_default_noise"] = gpd.log_likelihood(h1)
        # user-composed squared-exponential kernel (examples/SingleTaskTest.ipynb `skernel` pattern)
        def se(x1, x2, h):
            return h[0] * rk.squared_exponential_kernel(rk.get_distance_matrix(x1, x2), h[1])
        hs = np.array([1.2, 0.2])
        gps = fv.GP(x[:300], y[:300], init_hyperparameters=hs, noise_variances=noise[:300], kernel_function=se)
        out["se_hps"], out["lml_se"] = hs, gps.log_likelihood(hs * 1.1)
        save("dense_lml_" + tag, **out)


def multitask():
    rng = np.random.default_rng(3)
    n, T = 120, 3
    x = rng.random((n, 2))
    y = np.stack([np.sin((3 + t) * x[:, 0]) + np.cos(2 * x[:, 1]) + 0.1 * rng.standard_normal(n)
                  for t in range(T)], axis=1)
    y[5, 1] = np.nan
    y[17, 2] = np.nan
    noise = np.full((n, T), 1e-2)
    h = np.array([1.0, .3, .3, 2.0])
    gp = fv.fvGP(x, y, init_hyperparameters=h, noise_variances=noise)
    h1 = h * 1.05
    xp = rng.random((7, 2))
    pm = gp.posterior_mean(xp)
    save("multitask", x=x, y=y, noise=noise, h=h, h1=h1, x_index=gp.x_data, y_flat=gp.y_data,
         v_flat=gp.V, lml=gp.log_likelihood(h1), grad=gp.neg_log_likelihood_gradient(h1),
         x_pred=xp, post_mean=pm["m(x)"], post_mean_flat=pm["m(x)_flat"])


def csr_parts(K):
    K = K.tocsr()
    K.sort_indices()
    return dict(indptr=K.indptr, indices=K.indices, data=K.data)


def gp2scale():
    out = {}
    # tests/test_fvgp.py:1711 inputs (support-aware == dense); we store the dense (defining) block
    rs = np.random.RandomState(0)
    a, b = rs.rand(40, 3), rs.rand(30, 3)
    hp = np.array([1.7, .3, .4, .5])
    out["t1711_x1"], out["t1711_x2"], out["t1711_hps"] = a, b, hp
    out["t1711_dense"] = rk.wendland_anisotropic_gp2Scale_cpu(a, b, hp)
    rs = np.random.RandomState(1)
    a = rs.rand(25, 2)
    hp2 = np.array([2.5, .6, .4])
    out["t1724_x"], out["t1724_hps"] = a, hp2
    out["t1724_dense"] = rk.wendland_anisotropic_gp2Scale_cpu(a, a, hp2)
    # tests/test_fvgp.py:3027 distributed_covariance, blockwise B=10, symmetric and rectangular
    rng = np.random.default_rng(42)
    x1, x2 = rng.random((57, 2)), rng.random((23, 2))
    hp3 = np.array([2.0, .4, .35])
    cl = fv.__stub_client__()
    f1, f2 = cl.scatter(x1), cl.scatter(x2)
    Ks = g2.distributed_covariance(cl, rk.wendland_anisotropic_gp2Scale_cpu, hp3, f1, 57, f1, 57, 10,
                                   symmetric=True, distribution="blockwise")
    Kr = g2.distributed_covariance(cl, rk.wendland_anisotropic_gp2Scale_cpu, hp3, f1, 57, f2, 23, 10,
                                   symmetric=False, distribution="blockwise")
    out["t3027_x1"], out["t3027_x2"], out["t3027_hps"] = x1, x2, hp3
    for k, v in csr_parts(Ks).items():
        out["t3027_sym_" + k] = v
    for k, v in csr_parts(Kr).items():
        out["t3027_rect_" + k] = v
    save("gp2scale_blocks", **out)

    # full gp2Scale GP with the exact sparse-LU path: test_fvgp.py:3152 shape and a 3-D N=4000 case
    for tag, n, d, seed, hp, B in (("t3152", 90, 2, 11, np.array([1.2, .35, .3]), 25),
                                   ("c4small", 4000, 3, 4, np.array([1.0, .11, .12, .1]), 1000)):
        rng = np.random.default_rng(seed)
        x = rng.random((n, d))
        y = np.sin(8 * np.linalg.norm(x, axis=1)) + 0.1 * rng.standard_normal(n)
        noise = np.full(n, 1e-2)
        gp = fv.GP(x, y, init_hyperparameters=hp, noise_variances=noise, gp2Scale=True,
                   dask_client=fv.__stub_client__(), gp2Scale_batch_size=B, linalg_mode="sparseLU")
        h1 = hp * np.concatenate([[1.0], np.full(d, 1.05)])
        o = dict(x=x, y=y, noise=noise, h0=hp, h1=h1, lml_h0=gp.log_likelihood(hp), lml_h1=gp.log_likelihood(h1),
                 KVinvY_h0=gp.kv.KVinvY, logdet_h0=gp.kv.logdet_KV)
        o.update(csr_parts(gp.K))
        K1 = gp.prior.compute_prior_covariance_matrix(x, h1)
        for k, v in csr_parts(K1).items():
            o["h1_" + k] = v
        xp = np.random.default_rng(seed + 7).random((9, d))
        o["x_pred"] = xp
        o["post_mean"] = gp.posterior_mean(xp)["m(x)"]
        o["post_var"] = gp.posterior_covariance(xp)["v(x)"]
        if tag == "c4small":       # keep the fixture small: drop the bulky value arrays, keep a checksum
            for key in ("data", "h1_data"):
                o[key + "_sum"] = np.sum(o[key])
                o[key + "_head"] = o[key][:2000]
                del o[key]
        save("gp2scale_" + tag, **o)
        del gp


def robust_kernels():
    """SURVEY 8a rows a3 / a16 / a17 beyond the four basic radial names: the *_robust variants (kernels.py:36, :77,
    :144, :191), wendland_kernel (:336), the dense wendland_anisotropic (:355) and the support-aware sparse block
    kernel wendland_anisotropic_gp2Scale_cpu_sparse (:724)."""
    rng = np.random.default_rng(202)
    x1 = rng.random((41, 3)) * 1.5 - 0.25
    x2 = rng.random((33, 3)) * 1.5 - 0.25
    hps = np.array([1.4, 0.5, 0.7, 0.9])
    phi = 1.35
    out = dict(x1=x1, x2=x2, hps=hps, phi=np.array(phi))
    d_iso = rk.get_distance_matrix(x1, x2)
    d_ani = rk.get_anisotropic_distance_matrix(x1, x2, hps[1:])
    for nm, f in (("se", rk.squared_exponential_kernel_robust), ("exp", rk.exponential_kernel_robust),
                  ("matern32", rk.matern_kernel_diff1_robust), ("matern52", rk.matern_kernel_diff2_robust)):
        out[nm + "_robust_iso"] = f(d_iso, phi)
        out[nm + "_robust_ani"] = f(d_ani, phi)
    out["wendland_kernel_ani"] = rk.wendland_kernel(np.array(d_ani, copy=True))
    out["wendland_anisotropic_12"] = rk.wendland_anisotropic(x1, x2, hps)
    out["wendland_anisotropic_11"] = rk.wendland_anisotropic(x1, x1, hps)
    sp12 = rk.wendland_anisotropic_gp2Scale_cpu_sparse(x1, x2, hps)
    out["wendland_sparse_12"] = np.asarray(sp12.toarray())
    out["wendland_block_12"] = rk.wendland_anisotropic_gp2Scale_cpu(x1, x2, hps)
    save("robust_kernels", **out)


if __name__ == "__main__":
    dense_kernels()
    dense_lml()
    multitask()
    gp2scale()
    robust_kernels()
