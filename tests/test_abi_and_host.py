"""CPU-side checks: the C-ABI library loads and exports every symbol include/*.h declares, the
product path fails loudly without a GPU, and the host-side logic mirrors the reference."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "fvgp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fvgp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from fvgp_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fvgp_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    assert set(_lib.EXPORTED_SYMBOLS) == set(syms), set(_lib.EXPORTED_SYMBOLS) ^ set(syms)
    assert lib.fvgp_version() >= 100
    # size helpers are pure host functions: callable without a GPU
    assert lib.fvgp_chol_workspace_len(130) == 2 * 128 * 128
    assert lib.fvgp_wendland_aabb_len(100, 3) == (4 + 1) * 6
    assert lib.fvgp_bjacobi_len(33) == 2 * 1024
    assert lib.fvgp_potrs_work_len(1000) >= 2 * 1000 + 8 * 1000
    # gp2Scale work units: at most 32 column chunks, each a multiple of 32 super tiles (1024 tiles, 32768 points)
    assert lib.fvgp_wendland_chunk_len(10, 5000) == 10                      # one chunk
    assert lib.fvgp_wendland_chunk_len(7, 1_000_000) == 31 * 7             # 977 super tiles -> 31 chunks of 32
    assert lib.fvgp_wendland_chunk_len(1, 40_000_000) <= 32
    assert lib.fvgp_wendland_chunk_len(0, 100) == 1
    assert lib.fvgp_lanczos_work_len(1000, 20) >= 3 * 16 * 1000 + 2 * 20 * 16


def test_rounded_allocation_sizes_are_stable():
    """nnz-sized buffers are rounded so that a few per cent of drift maps to the same allocator block."""
    def cap(count):
        step = max(1 << 18, (1 << max(int(count), 1).bit_length() - 1) >> 3)
        return (int(count) + step - 1) // step * step
    assert cap(94_124_622) == cap(95_900_000) == cap(99_000_000)
    assert cap(94_124_622) <= 1.13 * 94_124_622
    assert all(cap(c) >= c for c in (0, 1, 3806, 2 ** 31 + 5))
    import inspect
    from fvgp_b200 import _lib
    src = inspect.getsource(_lib.dev_empty_rounded)
    assert "bit_length" in src and "1 << 18" in src                          # the helper above mirrors the product code


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fvgp_b200 import GP, NativeLibraryError
    with pytest.warns(UserWarning):
        with pytest.raises(NativeLibraryError):
            GP(np.random.rand(8, 2), np.random.rand(8), init_hyperparameters=np.ones(3))


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fvgp_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f


def test_fvgp_transform_matches_reference(golden):
    from fvgp_b200.fvgp import fvGP
    g = golden("multitask")
    xi, yf, vf = fvGP._transform_index_set(g["x"], g["y"], g["noise"])
    assert np.array_equal(xi, g["x_index"]) and np.array_equal(yf, g["y_flat"][:, 0]) and np.array_equal(vf, g["v_flat"])


def test_cartesian_product_is_task_major():
    from fvgp_b200.gp_posterior import GPposterior
    x = np.arange(6.0).reshape(3, 2)
    out = GPposterior.cartesian_product(x, np.array([0.0, 1.0]))
    ref = np.array([np.append(x[i], t) for t in (0.0, 1.0) for i in range(3)])      # gp_posterior.py:586-606
    assert np.array_equal(out, ref)


def test_lazy_kernel_algebra():
    from fvgp_b200 import kernels as K
    from fvgp_b200 import _lib as L
    x = np.random.rand(5, 2)
    d = K.get_distance_matrix(x, x)
    assert isinstance(d, K.Distance) and d.same and np.all(d.inv_scale == 1.0)
    r = 2.0 * K.squared_exponential_kernel(d, 0.3) * 1.5
    assert isinstance(r, K.Radial) and r.kind == L.K_SQEXP and r.amp == 3.0 and r.length == 0.3
    a = K.get_anisotropic_distance_matrix(x, np.random.rand(4, 2), np.array([.5, .25]))
    assert not a.same and np.allclose(a.inv_scale, [2.0, 4.0]) and a.shape == (5, 4)
    assert K.matern_kernel_diff1(a, 1.0).kind == L.K_MATERN32
    assert K.matern_kernel_diff2(a, 1.0).kind == L.K_MATERN52
    assert K.exponential_kernel(a, 1.0).kind == L.K_EXP
    w = K.wendland_anisotropic_gp2Scale_cpu(x, x, np.array([1.0, .1, .1]))
    assert isinstance(w, K.SparseWendland) and w.same
    # hyperparameters arrive as numpy scalars (hps[0] of an ndarray): numpy's operators run first and must hand the
    # product back to the lazy expression instead of converting it through __array__ (which would leave the fused path)
    h = np.array([1.3, 0.4, 0.6])
    r = h[0] * K.squared_exponential_kernel(d, h[1])
    assert isinstance(r, K.Radial) and r.amp == 1.3 and r.length == 0.4
    r = h[0] ** 2 * K.matern_kernel_diff2(K.get_anisotropic_distance_matrix(x, x, h[1:3]), 1.0) / np.float64(2.0)
    assert isinstance(r, K.Radial) and r.amp == 1.3 ** 2 / 2.0 and np.allclose(r.dist.inv_scale, 1.0 / h[1:3])
    assert isinstance(np.array(2.0) * r, K.Radial) and isinstance(np.multiply(r, h[0]), K.Radial)


def test_gp2scale_mode_thresholds():
    """gp_kv.py:182-188 (tests/test_fvgp.py:5110)."""
    from fvgp_b200.gp_kv import GPkv, resolve_gp2scale_linalg_mode

    class D:
        x_data = np.zeros((3000, 1))
        args = {}
    kv = GPkv.__new__(GPkv)
    kv.data, kv.linalg_mode = D(), None
    assert kv._set_gp2Scale_mode(int(0.00005 * 3000 ** 2)) == "sparseLU"
    assert kv._set_gp2Scale_mode(int(0.01 * 3000 ** 2)) == "sparseMINRES"
    D.x_data = np.zeros((1500, 1))
    assert kv._set_gp2Scale_mode(int(0.01 * 1500 ** 2)) == "Chol"
    kv.linalg_mode = "sparseCG"
    assert kv._set_gp2Scale_mode(10) == "sparseCG"
    assert resolve_gp2scale_linalg_mode("sparseCGpre_ilu", {})[0] == "sparseCGpre"
    assert resolve_gp2scale_linalg_mode("sparseCGpre_ilu", {})[1]["sparse_preconditioner_type"] == "ilu"


def test_addKV_formats():
    """gp_kv.py:640-669 (tests/test_fvgp.py:4231)."""
    import scipy.sparse as sp
    from fvgp_b200.gp_kv import GPkv
    K = np.arange(9.0).reshape(3, 3)
    V = np.array([.1, .2, .3])
    assert np.allclose(GPkv.addKV(K, V), K + np.diag(V))
    assert np.allclose(GPkv.addKV(K, np.diag(V)), K + np.diag(V))
    assert np.allclose(GPkv.addKV(sp.csr_matrix(K), V).toarray(), K + np.diag(V))


def test_training_drivers_on_a_toy_objective():
    from fvgp_b200.gp_training import GPtraining
    tr = GPtraining(None, np.array([0.5, 0.5]))
    target = np.array([0.3, 0.7])
    bounds = np.array([[0.0, 1.0], [0.0, 1.0]])

    def nll(h):
        return float(np.sum((h - target) ** 2))
    out = tr.train(objective_function=nll, objective_function_gradient=lambda h: 2 * (h - target),
                   hyperparameter_bounds=bounds, init_hyperparameters=np.array([.5, .5]), method="local", max_iter=50)
    assert np.allclose(out, target, atol=1e-4)
    out = tr.train(objective_function=nll, hyperparameter_bounds=bounds, init_hyperparameters=np.array([.5, .5]),
                   method="global", max_iter=40, pop_size=10)
    assert np.allclose(out, target, atol=1e-2)
    out = tr.train(objective_function=lambda h: -200 * nll(h), hyperparameter_bounds=bounds,
                   init_hyperparameters=np.array([.5, .5]), method="mcmc", max_iter=1500, mcmc_args={"seed": 1})
    assert np.allclose(out, target, atol=0.05) and tr.mcmc_info["acceptance rate"] > 0.05
    out = tr.train(objective_function=nll, objective_function_gradient=lambda h: 2 * (h - target),
                   hyperparameter_bounds=bounds, init_hyperparameters=np.array([.5, .5]), method="adam", max_iter=400)
    assert np.allclose(out, target, atol=2e-2)
    with pytest.raises(Exception):
        tr.train(objective_function=nll, hyperparameter_bounds=bounds, init_hyperparameters=np.array([2., 2.]))


def test_bench_reference_arm_prints_contract_line():
    import json
    import sys
    env = dict(os.environ, FVGP_REF_BUDGET_S="15")               # keeps the size ladder at N <= 1000 here
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-n", "400", "--no-c4"], capture_output=True, text=True,
                         timeout=300, env=env)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["value"] > 0
    # the unmodified reference when baseline/_ref is installed (build container, GPU box), else the oracle port
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "fvgp"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["fit"]["b"] >= 0 and "400" in line["measured_seconds_by_n"]


def test_speculative_mcmc_samples_the_same_distribution():
    """run_mcmc with a population evaluator (batches of proposals around the current state, consumed up to the first
    acceptance) must sample the same target as the one-proposal-at-a-time chain, with far fewer likelihood calls."""
    from fvgp_b200.gp_training import run_mcmc
    mu, sd = np.array([0.4, 1.2]), np.array([0.1, 0.3])

    def ll(t):
        return float(-0.5 * np.sum(((t - mu) / sd) ** 2))

    seen = []

    def pop(T):
        seen.append(len(T))
        return np.array([ll(t) for t in T])

    bounds = np.array([[-2.0, 3.0], [-2.0, 4.0]])
    stats = {}
    for name, fn in (("seq", None), ("pop", pop)):
        means, sds, calls = [], [], []
        for seed in range(4):
            r = run_mcmc(ll, np.zeros(2), bounds, n_updates=3000, seed=seed, population_log_likelihood=fn)
            assert len(r["x"]) == 3001 and len(r["f(x)"]) == 3001
            means.append(r["mean(x)"]); sds.append(np.sqrt(r["var(x)"])); calls.append(r["likelihood calls"])
        stats[name] = (np.mean(means, axis=0), np.mean(sds, axis=0), np.mean(calls))
    for name in ("seq", "pop"):
        assert np.allclose(stats[name][0], mu, atol=0.03) and np.allclose(stats[name][1], sd, rtol=0.15), stats[name]
    assert stats["seq"][2] == 3000 and stats["pop"][2] < 0.6 * 3000 and max(seen) <= 16 and min(seen) >= 2

    # a population call that fails (a speculative proposal that is not positive definite) falls back to one-by-one
    def bad_pop(T):
        raise RuntimeError("Linear algebra failed")
    r = run_mcmc(ll, np.zeros(2), bounds, n_updates=300, seed=0, population_log_likelihood=bad_pop)
    assert len(r["x"]) == 301 and r["likelihood calls"] >= 250


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype in include/fvgp_b200.h has as many parameters as its ctypes binding (guards ABI drift: a missing
    or extra argument in _lib._SIGNATURES would silently shift every later pointer)."""
    from fvgp_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "fvgp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = dict(re.findall(r"\b(fvgp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt))
    assert set(protos) == set(_lib._SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib._SIGNATURES[name][1]), (name, n, len(_lib._SIGNATURES[name][1]))
    # pointer-vs-scalar agreement for the population entry point (the widest signature)
    import ctypes
    sig = _lib._SIGNATURES["fvgp_lml_population"][1]
    decl = [p.strip() for p in protos["fvgp_lml_population"].split(",")]
    for ctype, d in zip(sig, decl):
        is_ptr = "*" in d
        assert is_ptr == (ctype is ctypes.c_void_p or hasattr(ctype, "contents") or ctype.__name__.startswith("LP_")), (d, ctype)


def test_install_as_fvgp_registers_the_reference_import_names():
    """`fvgp_b200.install_as_fvgp()`: unmodified `from fvgp import GP` / `from fvgp.kernels import ...` user code lands
    on this package (run in a subprocess: the reference must not already be imported under that name)."""
    import sys
    code = ("import sys; sys.path.insert(0, %r); import fvgp_b200; fvgp_b200.install_as_fvgp();"
            "from fvgp import GP, fvGP; from fvgp.kernels import squared_exponential_kernel, get_distance_matrix;"
            "import fvgp.gp_kv as kv; import inspect;"
            "assert GP is fvgp_b200.GP and kv.GPkv is fvgp_b200.gp_kv.GPkv;"
            "assert inspect.signature(GP.__init__).parameters['compute_device'].default == 'cpu';"
            "print('ok')") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.stdout.strip().endswith("ok"), out.stderr[-800:]
