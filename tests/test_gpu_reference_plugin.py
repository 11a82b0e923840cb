"""The ctypes stubs of INTEGRATION.md (integration/fvgp_reference_plugin.py), executed against the UNMODIFIED
reference package (baseline/_ref, installed by baseline/install_ref.py): the reference's own GP object drives our
C ABI through its operator seams, and its results equal those of its stock numpy / scipy path."""
import os
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")


@pytest.fixture(scope="module")
def ref():
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "fvgp")):
        pytest.skip("baseline/_ref is not installed (python baseline/install_ref.py where /root/reference exists)")
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import install_ref
    fv = install_ref.import_reference()
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import fvgp_reference_plugin as plug
    return fv, plug


def _data(n, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3))
    y = np.sin(5 * x[:, 0]) * np.cos(3 * x[:, 1]) + x[:, 2] + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2)


def test_reference_gp_through_the_linalg_and_kernel_seams(ref):
    fv, plug = ref
    x, y, noise = _data(900)
    h = np.array([1.1, .3, .45, .5])
    stock = fv.GP(x, y, init_hyperparameters=h, noise_variances=noise)
    ours = fv.GP(x, y, init_hyperparameters=h, noise_variances=noise, kernel_function=plug.b200_default_kernel,
                 linalg_mode=plug.b200_linalg_mode)
    h1 = h * 1.04
    assert abs(ours.log_likelihood(h1) / stock.log_likelihood(h1) - 1) <= 1e-10
    assert abs(ours.log_likelihood() / stock.log_likelihood() - 1) <= 1e-10
    xp = np.random.default_rng(1).random((30, 3))
    assert np.allclose(ours.posterior_mean(xp)["m(x)"], stock.posterior_mean(xp)["m(x)"], rtol=1e-9, atol=1e-11)
    assert np.allclose(ours.posterior_covariance(xp)["v(x)"], stock.posterior_covariance(xp)["v(x)"], rtol=1e-7, atol=1e-10)
    K_ours, K_stock = ours.prior.K, stock.prior.K
    assert np.max(np.abs(K_ours - K_stock) / np.abs(K_stock)) <= 1e-12


def test_reference_gp2scale_assembly_through_the_kernel_seam(ref):
    """distributed_covariance (gp2Scale_covariance.py:313-431) with OUR block kernel plugged into the reference's
    kernel seam: the assembled CSR equals the one its own dense block kernel produces, bit for bit in the pattern."""
    fv, plug = ref
    from fvgp import gp2Scale_covariance as g2s
    from fvgp import kernels as rk
    import ref_stubs
    x, _, _ = _data(1500, seed=3)
    th = np.array([1.2, .12, .11, .13])
    client = ref_stubs.Client()
    fut = client.scatter(x)
    args = dict(symmetric=True, distribution="blockwise", k_n_params=3, args={})
    K_stock = g2s.distributed_covariance(client, rk.wendland_anisotropic_gp2Scale_cpu, th, fut, len(x), fut, len(x), 400, **args)
    K_ours = g2s.distributed_covariance(client, plug.b200_wendland_gp2Scale, th, fut, len(x), fut, len(x), 400, **args)
    K_stock.sort_indices(), K_ours.sort_indices()
    assert np.array_equal(K_stock.indptr, K_ours.indptr) and np.array_equal(K_stock.indices, K_ours.indices)
    assert np.max(np.abs(K_stock.data - K_ours.data) / np.abs(K_stock.data)) <= 1e-12
