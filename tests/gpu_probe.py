"""One-shot GPU diagnostic: roofline denominators + correctness of every kernel against the
oracle, printing everything and never stopping at the first failure.  Run on the B200 box:

    python tests/gpu_probe.py [--quick] > gpurun_out/probe.log

(The pytest `-m gpu` suite is the gate; this script exists to get maximal information out of
one gpurun round trip while kernels are being brought up.)
"""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from fvgp_b200 import _lib as L  # noqa: E402
from fvgp_b200 import ops  # noqa: E402
from oracle import fvgp_oracle as orc  # noqa: E402

QUICK = "--quick" in sys.argv
RESULTS = {}
FAILS = []


ONLY = os.environ.get("PROBE_ONLY", "")


def section(name):
    def deco(fn):
        if ONLY and ONLY not in name:
            return fn
        print(f"\n===== {name} =====", flush=True)
        t0 = time.time()
        try:
            fn()
        except Exception:
            FAILS.append(name)
            print(f"[EXCEPTION in {name}]")
            traceback.print_exc()
        print(f"[{name}: {time.time() - t0:.1f}s]", flush=True)
        return fn
    return deco


def check(label, err, tol):
    ok = bool(err <= tol)
    print(f"  {'ok  ' if ok else 'FAIL'} {label}: err={err:.3e} tol={tol:.1e}", flush=True)
    if not ok:
        FAILS.append(label)
    return ok


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def cuda_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


@section("environment")
def _env():
    import multiprocessing
    print("gpu:", torch.cuda.get_device_name(0), "| cpus:", multiprocessing.cpu_count())
    with open("/proc/meminfo") as fh:
        print(fh.readline().strip())
    print("lib version:", L.load().fvgp_version())


@section("fp64 peaks")
def _peaks():
    import ctypes
    lib = L.load()
    scratch = L.dev_empty((148 * 8 * 256,))
    out = ctypes.c_double()
    for which, nm in ((0, "dmma"), (1, "dfma")):
        for cps in (1, 2, 4):
            lib.fvgp_bench_fp64_peak(which, cps, 20000, L.ptr(scratch), ctypes.byref(out), L.stream_ptr())
            print(f"  {nm} ctas/sm={cps}: {out.value:.2f} TFLOP/s")
            RESULTS[f"{nm}_peak_tflops_c{cps}"] = out.value
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    t = cuda_time(lambda: torch.matmul(a, b), reps=3)
    RESULTS["cublas_dgemm_tflops"] = 2 * n ** 3 / t / 1e12
    print(f"  cuBLAS dgemm {n}: {RESULTS['cublas_dgemm_tflops']:.2f} TFLOP/s")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    t = cuda_time(lambda: ops.dgemm_nt(a, b, c), reps=3)
    RESULTS["our_dgemm_tflops"] = 2 * n ** 3 / t / 1e12
    print(f"  our   dgemm_nt {n}: {RESULTS['our_dgemm_tflops']:.2f} TFLOP/s")
    ref = a @ b.T
    check("dgemm_nt 8192 vs cuBLAS", float((c - ref).abs().max() / ref.abs().max()), 1e-13)
    t = cuda_time(lambda: ops.dgemm_nt(a, a, c, lower=True), reps=3)
    print(f"  our   syrk-lower {n}: {n ** 3 / t / 1e12:.2f} TFLOP/s (useful flops n^3)")
    RESULTS["our_syrk_tflops"] = n ** 3 / t / 1e12
    del a, b, c, ref
    for n in ((8192,) if QUICK else (8192, 16384)):
        m = torch.randn(n, n, dtype=torch.float64, device="cuda")
        spd = m @ m.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
        del m
        t = cuda_time(lambda: torch.linalg.cholesky(spd), reps=2)
        RESULTS[f"cusolver_potrf_tflops_{n}"] = n ** 3 / 3 / t / 1e12
        print(f"  torch.linalg.cholesky {n}: {t * 1e3:.1f} ms = {n ** 3 / 3 / t / 1e12:.2f} TFLOP/s")
        buf, ld = L.dev_matrix(n, n)

        def ours():
            buf[:, :n] = spd
            return ops.potrf(buf, ld, n)
        tcopy = cuda_time(lambda: buf[:, :n].copy_(spd), reps=2)
        t = cuda_time(ours, reps=2) - tcopy
        RESULTS[f"our_potrf_tflops_{n}"] = n ** 3 / 3 / t / 1e12
        print(f"  our potrf {n}: {t * 1e3:.1f} ms = {n ** 3 / 3 / t / 1e12:.2f} TFLOP/s")
        f = ours()
        ref = torch.linalg.cholesky(spd)
        check(f"potrf {n} vs cusolver", float((f.lower() - ref).abs().max() / ref.abs().max()), 1e-12)
        t = cuda_time(lambda: (buf[:, :n].copy_(ref), ops.potri(ops.CholFactor(buf, ld, n, f.tileinv))), reps=1, warm=1) - tcopy
        print(f"  our potri {n}: {t * 1e3:.1f} ms = {2 * n ** 3 / 3 / t / 1e12:.2f} TFLOP/s")
        RESULTS[f"our_potri_tflops_{n}"] = 2 * n ** 3 / 3 / t / 1e12
        del spd, ref, buf, f


@section("dgemm variants / shapes")
def _gemm_shapes():
    rng = np.random.default_rng(0)
    for (m, n, k) in ((1, 64, 64), (130, 70, 33), (257, 129, 515), (64, 64, 64), (300, 300, 1000)):
        a, b = dev(rng.standard_normal((m, k + (k % 2)))), dev(rng.standard_normal((n, k + (k % 2))))
        c0 = rng.standard_normal((m, n + (n % 2)))
        c = dev(c0)
        ops.dgemm_nt(a[:, :k], b[:, :k], c[:, :n], alpha=-1.0, beta=1.0)
        ref = c0[:, :n] - a[:, :k].cpu().numpy() @ b[:, :k].cpu().numpy().T
        check(f"dgemm_nt {m}x{n}x{k} beta=1", float(np.abs(c[:, :n].cpu().numpy() - ref).max()), 1e-11)
        if n % 2 and c0.shape[1] > n:
            check("   padding column untouched", float(np.abs(c[:, n].cpu().numpy() - c0[:, n]).max()), 0.0)


def elementwise_relerr(a, b):
    """max_ij |a - b| / |b| over entries with b != 0 (plus the absolute error where b == 0)."""
    a, b = np.asarray(a), np.asarray(b)
    nz = b != 0
    e1 = float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz]))) if nz.any() else 0.0
    e0 = float(np.max(np.abs(a[~nz]))) if (~nz).any() else 0.0
    return max(e1, e0)


@section("dense K-fill vs oracle")
def _kfill():
    rng = np.random.default_rng(5)
    for n1, n2, d in ((37, 29, 3), (200, 200, 1), (333, 333, 3), (130, 257, 2), (64, 64, 5), (129, 129, 4)):
        x1 = rng.random((n1, d)) * 3 - 1
        x2 = x1 if n1 == n2 else rng.random((n2, d))
        hps = np.concatenate([[1.7], 0.2 + rng.random(d)])
        dx1, dx2 = dev(x1), dev(x2)
        ref = orc.default_kernel(x1, x2, hps)
        both = np.vstack([x1, x2])
        for bounds, tag in ((None, "exact-diff"), ((both.min(axis=0), both.max(axis=0)), "centred")):
            buf, ld = ops.kfill(L.K_MATERN32, dx1, dx2, hps[0], 1.0 / hps[1:], 1.0, bounds=bounds)
            err = elementwise_relerr(buf[:, :n2].cpu().numpy(), ref)
            check(f"default kernel full {n1}x{n2} d={d} {tag} (max elementwise rel)", err, 1e-12)
            if n1 == n2:
                noise = rng.random(n1) * 0.1
                refn = orc.add_kv(ref, noise)
                for bulk in (1, 0):
                    L.load().fvgp_set_bulk_store(bulk)
                    buf, ld = ops.kfill(L.K_MATERN32, dx1, dx1, hps[0], 1.0 / hps[1:], 1.0, noise=dev(noise),
                                        mode=L.FILL_SYMMETRIC, bounds=bounds)
                    check(f"  symmetric+noise bulk={bulk} {tag}", elementwise_relerr(buf[:, :n2].cpu().numpy(), refn), 1e-12)
                L.load().fvgp_set_bulk_store(1)
                buf, ld = ops.kfill(L.K_MATERN32, dx1, dx1, hps[0], 1.0 / hps[1:], 1.0, noise=dev(noise),
                                    mode=L.FILL_LOWER, bounds=bounds)
                check(f"  lower+noise {tag}", elementwise_relerr(np.tril(buf[:, :n2].cpu().numpy()), np.tril(refn)), 1e-12)
    x1, x2 = rng.random((150, 3)), rng.random((90, 3))
    dx1, dx2 = dev(x1), dev(x2)
    d_iso, d_ani = orc.distance_matrix(x1, x2), orc.anisotropic_distance_matrix(x1, x2, np.array([.3, .5, .9]))
    one = np.ones(3)
    check("distance iso", relerr(ops.kfill(L.K_DISTANCE, dx1, dx2, 1.0, one)[0][:, :90].cpu().numpy(), d_iso), 1e-13)
    check("distance aniso", relerr(ops.kfill(L.K_DISTANCE, dx1, dx2, 1.0, 1 / np.array([.3, .5, .9]))[0][:, :90].cpu().numpy(), d_ani), 1e-13)
    bb = (np.zeros(3), np.ones(3))
    check("distance iso centred", elementwise_relerr(ops.kfill(L.K_DISTANCE, dx1, dx2, 1.0, one, bounds=bb)[0][:, :90].cpu().numpy(), d_iso), 1e-12)
    for kind, nm in ((L.K_SQEXP, "se"), (L.K_EXP, "exp"), (L.K_MATERN32, "matern32"), (L.K_MATERN52, "matern52")):
        ref = 2.5 * orc.RADIAL[nm](d_iso, 0.37)
        for bounds, tag in ((None, "exact-diff"), (bb, "centred")):
            got = ops.kfill(kind, dx1, dx2, 2.5, one, 0.37, bounds=bounds)[0][:, :90].cpu().numpy()
            check(f"{nm} iso {tag} (max elementwise rel)", elementwise_relerr(got, ref), 1e-12)
    # far-apart / tiny length scale: underflow region and the CENTRED_LIMIT fallback
    xs = np.vstack([rng.random((70, 2)), rng.random((70, 2)) + 50.0])
    hs = np.array([1.0, .02, .03])
    ref = orc.default_kernel(xs, xs, hs)
    got = ops.kfill(L.K_MATERN32, dev(xs), dev(xs), hs[0], 1 / hs[1:], 1.0, bounds=(xs.min(0), xs.max(0)))[0][:, :140].cpu().numpy()
    check("small length scale (fallback to exact-diff): |got| where ref underflows", float(np.abs(got[ref < 1e-290]).max()), 1e-290)
    check("small length scale, elementwise rel where ref > 1e-290", elementwise_relerr(np.where(ref > 1e-290, got, 0), np.where(ref > 1e-290, ref, 0)), 1e-12)
    refw = orc.wendland_block(x1, x2, np.array([1.3, .3, .5, .9]))
    got = ops.kfill(L.K_WENDLAND, dx1, dx2, 1.3, 1 / np.array([.3, .5, .9]), 1.0)[0][:, :90].cpu().numpy()
    check("wendland dense (abs)", float(np.abs(got - refw).max()), 1e-14)


@section("dense factorisation vs scipy")
def _chol():
    rng = np.random.default_rng(7)
    for n in (5, 64, 65, 100, 128, 129, 200, 513, 1000, 2100):
        x = rng.random((n, 3))
        hps = np.array([1.2, .3, .4, .5])
        noise = np.full(n, 1e-2)
        KV = orc.add_kv(orc.default_kernel(x, x, hps), noise)
        c = orc.chol_factor(KV)
        buf, ld = ops.kfill(L.K_MATERN32, dev(x), dev(x), hps[0], 1 / hps[1:], 1.0, noise=dev(noise), mode=L.FILL_LOWER)
        f = ops.potrf(buf, ld, n)
        check(f"potrf n={n}", float(np.abs(f.lower().cpu().numpy() - np.tril(c)).max()), 1e-11)
        check("  logdet", abs(ops.chol_logdet(f) / orc.chol_logdet(c) - 1), 1e-12)
        y = rng.standard_normal((n, 1))
        rhs = dev(y.T.copy())
        ops.potrs(f, rhs)
        check("  potrs nrhs=1", relerr(rhs.cpu().numpy().T, orc.chol_solve(c, y)), 1e-8)
        Y = rng.standard_normal((n, 7))
        rhs = dev(Y.T.copy())
        ops.potrs(f, rhs)
        ref = orc.chol_solve(c, Y)
        check("  potrs nrhs=7 (gemm path)", float(np.abs(rhs.cpu().numpy().T - ref).max() / np.abs(ref).max()), 1e-10)
        alpha = dev(orc.chol_solve(c, y)[:, 0])
        ops.potri(f)
        Kinv = np.linalg.inv(KV)
        got = np.tril(f.buf[:, :n].cpu().numpy())
        check("  potri (lower)", float(np.abs(got - np.tril(Kinv)).max() / np.abs(Kinv).max()), 1e-10)
        tr = ops.kgrad_trace_matern32(dev(x), hps, f.buf, f.ld, alpha)
        dK = orc.default_kernel_gradient(x, x, hps)
        W = Kinv - np.outer(alpha.cpu().numpy(), alpha.cpu().numpy())
        ref = np.array([np.sum(W * dK[h]) for h in range(4)])
        check("  gradient traces", relerr(tr, ref), 1e-9)
    # non-PD detection
    bad = dev(np.array([[1.0, 2.0], [2.0, 1.0]]))
    buf, ld = L.dev_matrix(2, 2)
    buf[:, :2] = bad
    try:
        ops.potrf(buf, ld, 2)
        check("non-PD raises", 1.0, 0.0)
    except L.NonPositiveDefiniteError as e:
        check(f"non-PD raises (pivot {e.pivot})", abs(e.pivot - 2), 0)


@section("full dense LML + gradient vs golden")
def _lml():
    for tag in ("c1", "c2"):
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"dense_lml_{tag}.npz")))
        x, y, noise = g["x"], g["y"], g["noise"]
        n = len(x)
        for hk in ("h0", "h1"):
            hps = g[hk]
            dx = dev(x)
            buf, ld = ops.kfill(L.K_MATERN32, dx, dx, hps[0], 1 / hps[1:], 1.0, noise=dev(noise), mode=L.FILL_LOWER)
            f = ops.potrf(buf, ld, n)
            ym = dev((y - y.mean())[None, :])
            alpha = ops.potrs(f, ym.clone())
            lml = -0.5 * (ops.dot(ym, alpha) + ops.chol_logdet(f) + n * np.log(2 * np.pi))
            check(f"LML {tag} {hk}", abs(lml / g['lml_' + hk] - 1), 1e-10)
            ops.potri(f)
            grad = 0.5 * ops.kgrad_trace_matern32(dx, hps, f.buf, f.ld, alpha[0])
            check(f"grad {tag} {hk}", relerr(grad, g["grad_" + hk]), 1e-8)


@section("gp2Scale CSR vs golden / oracle")
def _sparse():
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "gp2scale_blocks.npz")))
    x1, x2, h = g["t3027_x1"], g["t3027_x2"], g["t3027_hps"]
    for a, b, pre in ((x1, x1, "t3027_sym_"), (x1, x2, "t3027_rect_")):
        da = dev(a)
        db = da if b is a else dev(b)
        K = ops.wendland_csr(da, db, h).to_scipy()
        check(pre + "indptr", float(np.abs(K.indptr - g[pre + "indptr"]).max()), 0)
        same = K.indices.shape == g[pre + "indices"].shape and np.array_equal(K.indices, g[pre + "indices"])
        check(pre + "indices", 0.0 if same else 1.0, 0)
        if same:
            check(pre + "data", relerr(K.data, g[pre + "data"]), 1e-12)
    for tag in ("t3152", "c4small"):
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"gp2scale_{tag}.npz")))
        x, y, noise = g["x"], g["y"], g["noise"]
        dx = dev(x)
        for pre, hk in (("", "h0"), ("h1_", "h1")):
            K = ops.wendland_csr(dx, dx, g[hk]).to_scipy()
            same = np.array_equal(K.indptr, g[pre + "indptr"]) and np.array_equal(K.indices, g[pre + "indices"])
            check(f"{tag} {hk} pattern bit-exact (nnz {K.nnz})", 0.0 if same else 1.0, 0)
            ref = orc.gp2scale_covariance(x, x, g[hk], batch=1000, symmetric=True)
            if same:
                check("   values vs oracle", relerr(K.data, ref.data), 1e-12)
        KV = ops.wendland_csr(dx, dx, g["h0"], noise=dev(noise))
        ym = (y - y.mean())
        v = dev(np.random.default_rng(0).standard_normal(len(x)))
        ref = orc.add_kv(orc.gp2scale_covariance(x, x, g["h0"], batch=1000, symmetric=True), noise)
        rv = ref @ v.cpu().numpy()
        check("   spmv", float(np.abs(ops.spmv(KV, v).cpu().numpy() - rv).max() / np.abs(rv).max()), 1e-13)
        for pc in (None, "bjacobi"):
            M = ops.bjacobi(KV) if pc else None
            sol, info, iters, rr = ops.pcg(KV, dev(ym), rtol=1e-10, precond=M)
            print(f"   pcg precond={pc}: info={info} iters={iters} relres={rr:.2e}")
            check(f"   pcg({pc}) solution vs sparseLU", float(np.abs(sol.cpu().numpy() - g["KVinvY_h0"][:, 0]).max()
                                                             / np.abs(g["KVinvY_h0"]).max()), 1e-7)
        _, ref_it = orc.sparse_cg(ref, ym[:, None], rtol=1e-10)
        print(f"   scipy cg iterations: {ref_it}")
        est, var, _ = ops.slq_logdet(KV, degree=30, probes=40, seed=1)
        print(f"   slq logdet {est:.4f} +- {np.sqrt(var):.4f}  exact {g['logdet_h0']:.4f}")
        check("   slq within 5 sigma + 1%", abs(est - g["logdet_h0"]), 5 * np.sqrt(var) + 0.01 * abs(g["logdet_h0"]))
    # empty / disjoint / ragged
    K = ops.wendland_csr(dev(np.zeros((5, 2))), dev(np.ones((4, 2)) * 10), h)
    check("disjoint sets -> nnz 0", K.nnz, 0)


@section("timings")
def _timings():
    rng = np.random.default_rng(2)
    n = 8192 if QUICK else 30000
    x = dev(rng.random((n, 3)))
    noise = dev(np.full(n, 1e-2))
    hps = np.array([1.0, .3, .4, .5])
    out = L.dev_matrix(n, n)
    for mode, nm, bytes_ in ((L.FILL_SYMMETRIC, "symmetric", 8 * n * n), (L.FILL_FULL, "full", 8 * n * n),
                             (L.FILL_LOWER, "lower", 4 * n * n)):
        for bulk in ((1, 0) if mode == L.FILL_SYMMETRIC else (1,)):
            L.load().fvgp_set_bulk_store(bulk)
            for bounds, tag in ((None, "exact"), ((np.zeros(3), np.ones(3)), "centred")):
                t = cuda_time(lambda: ops.kfill(L.K_MATERN32, x, x, 1.0, 1 / hps[1:], 1.0, noise=noise, mode=mode, out=out,
                                                bounds=bounds), reps=5)
                print(f"  kfill {nm} {tag} bulk={bulk} n={n}: {t * 1e3:.2f} ms -> {bytes_ / t / 1e9:.0f} GB/s")
                RESULTS[f"kfill_{nm}_{tag}_bulk{bulk}_gbs"] = bytes_ / t / 1e9
    L.load().fvgp_set_bulk_store(1)
    x1d = dev(rng.random((n, 1)))
    for mode, nm in ((L.FILL_SYMMETRIC, "symmetric"), (L.FILL_FULL, "full")):
        t = cuda_time(lambda: ops.kfill(L.K_DISTANCE, x1d, x1d, 1.0, np.ones(1), 1.0, mode=mode, out=out), reps=5)
        print(f"  store-path ceiling (1-D distance fill, {nm}): {8 * n * n / t / 1e9:.0f} GB/s")
        t = cuda_time(lambda: ops.kfill(L.K_SQEXP, x, x, 1.0, 1 / hps[1:], 0.5, mode=mode, out=out,
                                        bounds=(np.zeros(3), np.ones(3))), reps=5)
        print(f"  squared-exponential fill ({nm}): {8 * n * n / t / 1e9:.0f} GB/s")
    t = cuda_time(lambda: out[0].fill_(1.0), reps=5)
    print(f"  torch fill_ same buffer: {out[0].numel() * 8 / t / 1e9:.0f} GB/s (write-only reference)")
    RESULTS["torch_fill_gbs"] = out[0].numel() * 8 / t / 1e9
    # end-to-end LML + gradient
    def lml_grad():
        buf, ld = ops.kfill(L.K_MATERN32, x, x, 1.0, 1 / hps[1:], 1.0, noise=noise, mode=L.FILL_LOWER, out=out)
        f = ops.potrf(buf, ld, n)
        ym = torch.ones(1, n, dtype=torch.float64, device="cuda")
        a = ops.potrs(f, ym.clone())
        ld_ = ops.chol_logdet(f)
        ops.potri(f)
        return ops.kgrad_trace_matern32(x, hps, f.buf, f.ld, a[0]), ld_
    t0 = time.time()
    lml_grad()
    torch.cuda.synchronize()
    t = time.time() - t0
    print(f"  LML+grad n={n}: {t:.3f} s  ({n ** 3 / t / 1e12:.2f} TFLOP/s on N^3)")
    RESULTS[f"lml_grad_seconds_{n}"] = t
    # sparse at scale
    ns = 100000 if QUICK else 400000
    xs = rng.random((ns, 3))
    key = np.lexsort((xs[:, 2] // 0.05, xs[:, 1] // 0.05, xs[:, 0] // 0.05))   # crude locality ordering
    xs = dev(xs[key])
    th = np.array([1.0, .029, .029, .029]) * np.array([1, 1, 1, 1]) * (1e6 / ns) ** (1 / 3)
    th[0] = 1.0
    t = cuda_time(lambda: ops.wendland_csr(xs, xs, th), reps=2)
    K = ops.wendland_csr(xs, xs, th, noise=dev(np.full(ns, 1e-2)))
    print(f"  wendland csr n={ns}: {t * 1e3:.1f} ms, nnz={K.nnz} ({K.nnz / ns:.1f}/row) -> {(12 * K.nnz) / t / 1e9:.1f} GB/s algorithmic")
    v = dev(rng.standard_normal(ns))
    yv = L.dev_empty((ns,))
    t = cuda_time(lambda: ops.spmv(K, v, yv), reps=5)
    print(f"  spmv: {t * 1e6:.0f} us -> {(12 * K.nnz + 16 * ns) / t / 1e9:.0f} GB/s")
    RESULTS["spmv_gbs"] = (12 * K.nnz + 16 * ns) / t / 1e9
    for pc in (None, "bjacobi"):
        M = ops.bjacobi(K) if pc else None
        t0 = time.time()
        sol, info, iters, rr = ops.pcg(K, v, rtol=1e-5, precond=M)
        torch.cuda.synchronize()
        print(f"  pcg({pc}) rtol 1e-5: {time.time() - t0:.3f}s iters={iters} info={info} relres={rr:.2e}")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    print("\n===== SUMMARY =====")
    print("failures:", FAILS if FAILS else "none")
    print(json.dumps(RESULTS, indent=1))
    with open(os.path.join(ROOT, "gpurun_out", f"probe_results_{ONLY.split()[0] if ONLY else 'all'}.json"), "w") as fh:
        json.dump({"results": RESULTS, "failures": FAILS}, fh, indent=1)
