"""CPU oracle for the fvGP training hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy/scipy restatement of the reference algorithm for every row of SURVEY.md
section 8(a).  Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline / reference
arm of `bench.py` may import this module; the product (`fvgp_b200/`) never does and
fails loudly when its CUDA library is missing.

Pinning: `tests/golden/make_golden.py` runs the UNMODIFIED reference (imported from
/root/reference through `tests/golden/ref_shim.py`) on seeded inputs and commits the
outputs under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function
below against those vectors (bit-exact for the gp2Scale pattern and values, <=1e-13
relative for dense K, <=1e-10 relative for LML / gradient).  Parity is therefore PINNED
for the dense path and for gp2Scale with an exact log-determinant.  The stochastic
(imate SLQ) log-determinant is un-vendored third-party arithmetic with no value-pinning
test in the reference: "parity unpinned" for that one quantity; it is checked against
the exact log-determinant instead.

All citations are file:line into /root/reference/fvgp/.
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SQRT3 = np.sqrt(3.0)
SQRT5 = np.sqrt(5.0)


# --------------------------------------------------------------------------- a2
def distance_matrix(x1, x2):
    """Isotropic Euclidean distance, kernels.py:440-458 (square, accumulate per axis, sqrt)."""
    acc = np.zeros((len(x1), len(x2)))
    for i in range(x1.shape[1]):
        acc += (x1[:, i][:, None] - x2[:, i][None, :]) ** 2
    return np.sqrt(acc)


def anisotropic_distance_matrix(x1, x2, scales):
    """Axis-scaled distance, kernels.py:461-481: |dx_i / scales[i]|**2 summed in axis order."""
    acc = np.zeros((len(x1), len(x2)))
    for i in range(x1.shape[1]):
        acc += np.abs(np.subtract.outer(x1[:, i], x2[:, i]) / scales[i]) ** 2
    return np.sqrt(acc)


# --------------------------------------------------------------------------- a3
def squared_exponential(d, length):
    """kernels.py:16-33."""
    return np.exp(-(d ** 2) / (2.0 * (length ** 2)))


def exponential(d, length):
    """kernels.py:56-74."""
    return np.exp(-d / length)


def matern32(d, length):
    """kernels.py:98-118 (matern_kernel_diff1)."""
    a = (SQRT3 * d) / length
    return (1.0 + a) * np.exp(-a)


def matern52(d, length):
    """kernels.py:166-188 (matern_kernel_diff2)."""
    return (1.0 + (SQRT5 * d) / length + (5.0 * d ** 2) / (3.0 * length ** 2)) * np.exp(-(SQRT5 * d) / length)


RADIAL = {"se": squared_exponential, "exp": exponential, "matern32": matern32, "matern52": matern52}


def squared_exponential_robust(d, phi):
    """kernels.py:36-53."""
    return np.exp(-(d ** 2) * (phi ** 2))


def exponential_robust(d, phi):
    """kernels.py:77-95."""
    return np.exp(-d * (phi ** 2))


def matern32_robust(d, phi):
    """kernels.py:144-163: 1/length -> phi**2."""
    return (1.0 + ((SQRT3 * d) * (phi ** 2))) * np.exp(-(SQRT3 * d) * (phi ** 2))


def matern52_robust(d, phi):
    """kernels.py:191-213.  The reference's quadratic term is (5 d^2)(3 phi^4) = 15 d^2 phi^4 (not 5/3): restated as is."""
    return (1.0 + ((SQRT5 * d) * (phi ** 2)) + ((5.0 * d ** 2) * (3.0 * phi ** 4))) * np.exp(-(SQRT5 * d) * (phi ** 2))


RADIAL_ROBUST = {"se": squared_exponential_robust, "exp": exponential_robust, "matern32": matern32_robust,
                 "matern52": matern52_robust}


def wendland(d):
    """kernels.py:336-352 on a copy (the reference clamps its argument in place)."""
    d = np.minimum(d, 1.0)
    return (1.0 - d) ** 8 * (32.0 * d ** 3 + 25.0 * d ** 2 + 8.0 * d + 1.0)


def wendland_anisotropic(x1, x2, hps):
    """kernels.py:355-378: dense anisotropic Wendland, hps = (amplitude, support radii)."""
    return hps[0] * wendland(anisotropic_distance_matrix(x1, x2, hps[1:]))


# --------------------------------------------------------------------------- a1
def default_kernel(x1, x2, hps):
    """ARD Matern-3/2, gp_prior.py:376-400: hps[0] * matern32(aniso distance with hps[1:], 1)."""
    return hps[0] * matern32(anisotropic_distance_matrix(x1, x2, hps[1:]), 1.0)


# --------------------------------------------------------------------------- a4
def default_kernel_gradient(x1, x2, hps):
    """d default_kernel / d hps, shape (H,U,V); gp_prior.py:421-436 with kernels.py:121-141.

    The length-scale derivative is evaluated in the reference's own form
    dadl*ea - (1+a)*dadl*ea with dadl = sqrt3 * (-dx_i^2 / (hps_i^3 d)), zero where d == 0.
    """
    d = anisotropic_distance_matrix(x1, x2, hps[1:])
    out = np.zeros((len(hps), len(x1), len(x2)))
    nz = d != 0.0
    a = SQRT3 * d
    ea = np.exp(-a)
    for i in range(x1.shape[1]):
        dd = np.zeros_like(d)
        dx2 = np.abs(np.subtract.outer(x1[:, i], x2[:, i])) ** 2
        dd[nz] = -dx2[nz] / (hps[1 + i] ** 3 * d[nz])
        dadl = SQRT3 * dd
        out[1 + i] = hps[0] * (dadl * ea - (1.0 + a) * dadl * ea)
    out[0] = matern32(d, 1.0)
    return out


# --------------------------------------------------------------------------- a6 / a7
def default_noise(y):
    """gp_likelihood.py:102-104: (mean(|y|)/100)^2 for every point."""
    return np.full(len(y), (np.mean(np.abs(y)) / 100.0) ** 2)


def add_kv(K, V):
    """gp_kv.py:640-669: K + diag(V) for a vector V (dense copy+fill_diagonal, sparse setdiag)."""
    if sp.issparse(K):
        KV = K.copy().tocsr()
        KV.setdiag(K.diagonal() + V)
        return KV
    KV = K.copy()
    np.fill_diagonal(KV, np.diag(K) + V)
    return KV


# --------------------------------------------------------------------------- a8-a10
class NonPositiveDefinite(Exception):
    """Mirrors NonPositiveDefiniteError, gp_lin_alg.py:27-58."""


def chol_factor(KV):
    """gp_lin_alg.py:237-269: scipy cho_factor(lower=True)."""
    try:
        c, _ = sla.cho_factor(KV, lower=True)
    except np.linalg.LinAlgError as exc:
        raise NonPositiveDefinite(str(exc)) from exc
    return c


def chol_solve(c, b):
    """gp_lin_alg.py:289-328."""
    return sla.cho_solve((c, True), b)


def chol_logdet(c):
    """gp_lin_alg.py:331-360: 2 * sum(log|diag|)."""
    return 2.0 * np.sum(np.log(np.abs(np.diag(c))))


# --------------------------------------------------------------------------- a12
def log_likelihood_from(KVinvY, logdet, y_minus_m):
    """gp_marginal_likelihood.py:171-178; the quadratic form is averaged over y columns."""
    n, r = y_minus_m.shape
    l1 = np.sum(y_minus_m * KVinvY) / r
    return -0.5 * (l1 + logdet + n * np.log(2.0 * np.pi))


def dense_log_likelihood(x, y, hps, noise=None, kernel=default_kernel, mean=None):
    """Full dense LML: K-fill, +V, Cholesky, solve, logdet."""
    y = y.reshape(len(y), -1)
    K = kernel(x, x, hps)
    V = default_noise(y) if noise is None else noise
    m = np.full(len(x), np.mean(y)) if mean is None else mean       # gp_prior.py:449-458
    c = chol_factor(add_kv(K, V))
    ym = y - m[:, None]
    return log_likelihood_from(chol_solve(c, ym), chol_logdet(c), ym)


# --------------------------------------------------------------------------- a13
def dense_neg_log_likelihood_gradient(x, y, hps, noise=None, component=0,
                                      kernel=default_kernel, kernel_grad=default_kernel_gradient,
                                      mean=None, dm_dh=None, economical=False):
    """Gradient of -LML, gp_marginal_likelihood.py:224-309.

    economical=False follows the reference literally: stacked LU solves of KV against
    dK/dh (gp_lin_alg.py:1581-1626), trace of each.  economical=True is the algebraically
    identical tr(KV^-1 dK) = sum(KV^-1 o dK) used for larger N on the CPU baseline.
    The `dL_dHm[i] == 0.0` switch of :301-308 is kept.
    """
    y = y.reshape(len(y), -1)
    n, H = len(x), len(hps)
    K = kernel(x, x, hps)
    V = default_noise(y) if noise is None else noise
    m = np.full(n, np.mean(y)) if mean is None else mean
    KV = add_kv(K, V)
    c = chol_factor(KV)
    b = chol_solve(c, y - m[:, None])[:, component]
    dK = kernel_grad(x, x, hps)                     # noise derivative is zero (gp_likelihood.py:112-120)
    dm = np.zeros((H, n)) if dm_dh is None else dm_dh
    grad = np.zeros(H)
    if economical:
        KVinv = chol_solve(c, np.eye(n))
    for i in range(H):
        gm = -dm[i] @ b
        if gm == 0.0:
            quad = b @ dK[i] @ b
            tr = np.sum(KVinv * dK[i]) if economical else np.trace(np.linalg.solve(KV, dK[i]))
            grad[i] = -0.5 * (quad - tr)
        grad[i] += gm
    return grad


# --------------------------------------------------------------------------- a16
def wendland_block(x1, x2, hps):
    """Dense compact-support block, kernels.py:502-528 (the DEFINING gp2Scale oracle).

    s accumulates ((x1_i - x2_i) / hps[1+i])**2 in axis order, each numpy ufunc rounding
    separately; d = min(1, sqrt(s)); value = hps[0] * (1-d)**8 * (32 d**3 + 25 d**2 + 8 d + 1).
    """
    s = np.zeros((len(x1), len(x2)))
    for i in range(x1.shape[1]):
        s += (np.subtract.outer(x1[:, i], x2[:, i]) / hps[1 + i]) ** 2
    d = np.sqrt(s)
    d[d > 1.0] = 1.0
    return hps[0] * (1.0 - d) ** 8 * (32.0 * d ** 3 + 25.0 * d ** 2 + 8.0 * d + 1.0)


# --------------------------------------------------------------------------- a18 / a19
def chunk_ranges(n, batch):
    """gp2Scale_covariance.py:48-61: chunks of at most `batch`, remainder last."""
    batch = max(1, int(batch))
    return [(s, min(s + batch, n)) for s in range(0, n, batch)]


def gp2scale_covariance(x1, x2, hps, batch=10000, kernel=wendland_block, symmetric=None, threads=1):
    """Blockwise sparse assembly, gp2Scale_covariance.py:136-170, 240-287, 313-431.

    Per block: dense kernel, np.nonzero pattern; symmetric diagonal blocks keep row<=col;
    off-diagonal entries are mirrored; COO -> canonical CSR (sorted int32 indices).
    threads > 1 evaluates the independent blocks in a thread pool (the reference maps them over dask workers);
    the triplets are concatenated in the same block order either way.
    """
    if symmetric is None:
        symmetric = x1 is x2
    n1, n2 = len(x1), len(x2)

    def block(ij):
        (i0, i1), (j0, j1) = ij
        blk = np.asarray(kernel(x1[i0:i1], x2[j0:j1], hps))
        r, c = np.nonzero(blk)
        v = blk[r, c]
        if symmetric and i0 == j0:
            keep = r <= c
            r, c, v = r[keep], c[keep], v[keep]
        r = r + i0
        c = c + j0
        out = [(r, c, v)]
        if symmetric:
            off = r != c
            out.append((c[off], r[off], v[off]))
        return out
    todo = [(ri, cj) for ri in chunk_ranges(n1, batch) for cj in chunk_ranges(n2, batch)
            if not (symmetric and ri[0] > cj[0])]
    if threads > 1 and len(todo) > 1:
        import concurrent.futures as cf
        with cf.ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(block, todo))
    else:
        parts = [block(t) for t in todo]
    rows = [t[0] for part in parts for t in part]
    cols = [t[1] for part in parts for t in part]
    vals = [t[2] for part in parts for t in part]
    if not rows:
        return sp.csr_matrix((n1, n2))
    idx = np.int32 if max(n1, n2) < 2 ** 31 else np.int64          # :107-114
    K = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows).astype(idx),
                                              np.concatenate(cols).astype(idx))), shape=(n1, n2))
    return K.tocsr()


# --------------------------------------------------------------------------- a20
def sparse_cg(KV, b, rtol=1e-5, x0=None, maxiter=None, M=None):
    """gp_lin_alg.py:1213-1291: scipy cg per right-hand-side column, atol=0."""
    b = b.reshape(len(b), -1)
    out = np.empty_like(b, dtype=float)
    iters = []
    for c in range(b.shape[1]):
        count = [0]
        sol, _ = spla.cg(KV, b[:, c], x0=None if x0 is None else x0[:, c], rtol=rtol, atol=0.0,
                         maxiter=maxiter, M=M, callback=lambda _x: count.__setitem__(0, count[0] + 1))
        out[:, c] = sol
        iters.append(count[0])
    return out, iters


def sparse_lu_solve_logdet(KV, b):
    """Exact sparse reference (mode sparseLU), gp_lin_alg.py:203-230 + LU logdet."""
    lu = spla.splu(KV.tocsc())
    logdet = float(np.sum(np.log(np.abs(lu.L.diagonal()))) + np.sum(np.log(np.abs(lu.U.diagonal()))))
    return lu.solve(b), logdet


def gp2scale_log_likelihood(x, y, hps, noise, batch=10000, exact=True, rtol=1e-10):
    """gp2Scale LML with the exact sparse-LU logdet."""
    y = y.reshape(len(y), -1)
    K = gp2scale_covariance(x, x, hps, batch=batch, symmetric=True)
    KV = add_kv(K, noise)
    m = np.full(len(x), np.mean(y))
    ym = y - m[:, None]
    alpha, logdet = sparse_lu_solve_logdet(KV, ym)
    if not exact:
        alpha, _ = sparse_cg(KV, ym, rtol=rtol)
    return log_likelihood_from(alpha.reshape(ym.shape), logdet, ym)


# --------------------------------------------------------------------------- a23
def fvgp_transform(x, y, noise=None):
    """fvGP index-set transform, fvgp.py:626-660: task-major stacking, NaN y dropped."""
    n, tasks = y.shape
    xs, ys, vs = [], [], []
    for t in range(tasks):
        keep = ~np.isnan(y[:, t])
        xs.append(np.column_stack([x[keep], np.full(int(keep.sum()), float(t))]))
        ys.append(y[keep, t])
        if noise is not None:
            vs.append(noise[keep, t])
    return np.vstack(xs), np.concatenate(ys), (np.concatenate(vs) if noise is not None else None)


# --------------------------------------------------------------------------- posterior (8f #1)
def posterior_mean_cov(x, y, hps, noise, x_pred, kernel=default_kernel):
    """gp_posterior.py:139-182, 229-288 (dense): mean = k^T KV^-1 (y-m) + m, S = kk - k^T KV^-1 k."""
    y = y.reshape(len(y), -1)
    m0 = np.mean(y)
    c = chol_factor(add_kv(kernel(x, x, hps), noise))
    k = kernel(x, x_pred, hps)
    mean = k.T @ chol_solve(c, y - m0) + m0
    S = kernel(x_pred, x_pred, hps) - k.T @ chol_solve(c, k)
    return mean[:, 0], S


# --------------------------------------------------------------------------- large-N variants (memory-lean)
# Same arithmetic per entry as the functions above (they CALL default_kernel / default_kernel_gradient on row
# blocks); only the storage differs: one N x N array, upper triangle, factorised / inverted in place by LAPACK, so
# that the oracle reaches the benchmarked sizes on the GPU box's host (N = 50 000 LML: 20 GB; the literal
# formulation needs (3H+3) N^2 doubles for the gradient).  Pinned against the plain functions in
# tests/test_oracle_golden.py, which are pinned against the reference's outputs.
def _fill_rows(x1, x2, hps, noise_diag=None, upper_only=False, block=1024, threads=None):
    """C-ordered block default_kernel(x1, x2, hps) evaluated by row blocks in a thread pool (gp_prior.py:376-400);
    noise_diag is added on the diagonal (gp_kv.py:640-669); upper_only skips the columns left of the diagonal."""
    import concurrent.futures as cf
    import os
    out = np.empty((len(x1), len(x2)))

    def work(r0):
        r1 = min(r0 + block, len(x1))
        c0 = r0 if upper_only else 0
        out[r0:r1, c0:] = default_kernel(x1[r0:r1], x2[c0:], hps)
        if noise_diag is not None:
            idx = np.arange(r0, r1)
            out[idx, idx] += noise_diag[r0:r1]
    with cf.ThreadPoolExecutor(threads or os.cpu_count()) as ex:
        list(ex.map(work, range(0, len(x1), block)))
    return out


def _dpotrf_lower_of_transpose(B):
    """In-place LAPACK dpotrf (what scipy cho_factor calls, gp_lin_alg.py:237-269) on the Fortran view B.T of a
    C-ordered array whose UPPER triangle is filled; returns the factor as that view (lower triangle)."""
    from scipy.linalg import lapack
    c, info = lapack.dpotrf(B.T, lower=1, overwrite_a=1, clean=0)
    if info != 0:
        raise NonPositiveDefinite(f"dpotrf info = {info}")
    assert np.shares_memory(c, B), "LAPACK copied the matrix"
    return c


def dense_log_likelihood_blocked(x, y, hps, noise, block=1024, threads=None, return_factor=False, split=None):
    """dense_log_likelihood for large N: blocked K-fill, in-place LAPACK dpotrf, triangular solves, 2 sum log diag
    (gp_lin_alg.py:289-360), LML (gp_marginal_likelihood.py:171-178).

    N^2 < 2^31: one dpotrf / dpotrs on the whole matrix.  Beyond (N >= 46 341; `split` forces it): ONE level of the
    2 x 2 block factorisation built from the same LAPACK / BLAS-3 routines on the half-size blocks
        L11 = chol(K11),  X = L11^-1 K12 (= L21^T),  L22 = chol(K22 - X^T X)
    because dpotrf on the whole N = 50 000 matrix crashed intermittently inside the threaded OpenBLAS of the GPU box
    (element offsets past 2^31: 2 of 3 runs)."""
    from scipy.linalg import blas, lapack
    y = y.reshape(len(y), -1)
    n = len(x)
    hps = np.asarray(hps, dtype=float)
    m = np.full(n, np.mean(y))
    ym = y - m[:, None]
    split = (n * n >= 2 ** 31) if split is None else split
    if not split:
        A = _fill_rows(x, x, hps, noise, upper_only=True, block=block, threads=threads)
        c = _dpotrf_lower_of_transpose(A)
        alpha, info = lapack.dpotrs(c, ym, lower=1)
        logdet = 2.0 * np.sum(np.log(np.abs(np.diagonal(c))))
        lml = log_likelihood_from(alpha, logdet, ym)
        return (lml, c, alpha) if return_factor else lml
    assert not return_factor, "the split factorisation does not return one factor array"
    h = n // 2
    L11 = _dpotrf_lower_of_transpose(_fill_rows(x[:h], x[:h], hps, noise[:h], upper_only=True, block=block, threads=threads))
    X = _fill_rows(x[:h], x[h:], hps, block=block, threads=threads)                      # K12, h x (n - h)
    X = sla.solve_triangular(L11, X, lower=True, overwrite_b=True, check_finite=False)   # L11^-1 K12 = L21^T
    B22 = _fill_rows(x[h:], x[h:], hps, noise[h:], upper_only=True, block=block, threads=threads)
    c22 = blas.dsyrk(-1.0, X, beta=1.0, c=B22.T, trans=1, lower=1, overwrite_c=1)        # lower(B22^T) -= X^T X
    assert np.shares_memory(c22, B22), "BLAS copied the matrix"
    L22 = _dpotrf_lower_of_transpose(B22)
    z1 = sla.solve_triangular(L11, ym[:h], lower=True, check_finite=False)
    z2 = sla.solve_triangular(L22, ym[h:] - X.T @ z1, lower=True, check_finite=False)
    a2 = sla.solve_triangular(L22, z2, lower=True, trans="T", check_finite=False)
    a1 = sla.solve_triangular(L11, z1 - X @ a2, lower=True, trans="T", check_finite=False)
    alpha = np.vstack([a1, a2])
    logdet = 2.0 * (np.sum(np.log(np.abs(np.diagonal(L11)))) + np.sum(np.log(np.abs(np.diagonal(L22)))))
    return log_likelihood_from(alpha, logdet, ym)


def dense_neg_log_likelihood_gradient_blocked(x, y, hps, noise, component=0, block=512, threads=None):
    """dense_neg_log_likelihood_gradient(economical=True) for large N: KV^-1 by LAPACK dpotri in place, traces
    sum(KV^-1 o dK_i) and b^T dK_i b accumulated over row blocks of the upper triangle (off-diagonal entries count
    twice), dK from default_kernel_gradient (gp_prior.py:421-436).  Returns (lml, grad)."""
    import concurrent.futures as cf
    import os
    from scipy.linalg import lapack
    hps = np.asarray(hps, dtype=float)
    n, H = len(x), len(hps)
    lml, c, alpha = dense_log_likelihood_blocked(x, y, hps, noise, 1024, threads, return_factor=True)
    b = alpha[:, component].copy()
    inv, info = lapack.dpotri(c, lower=1, overwrite_c=1)
    assert info == 0 and np.shares_memory(inv, c)
    W = inv.T                                                            # C view: W[r, c>=r] = KV^-1[r, c]

    def work(r0):
        r1 = min(r0 + block, n)
        dK = default_kernel_gradient(x[r0:r1], x[r0:], hps)             # (H, rows, n - r0)
        w = np.full((r1 - r0, n - r0), 2.0)
        k = np.arange(r1 - r0)
        w[k, k] = 1.0
        w[np.tril_indices(r1 - r0, -1)] = 0.0                            # inside the diagonal block: upper part only
        core = w * (np.outer(b[r0:r1], b[r0:]) - W[r0:r1, r0:])          # (b b^T - KV^-1) o weights
        return np.array([np.sum(core * dK[i]) for i in range(H)])
    with cf.ThreadPoolExecutor(threads or os.cpu_count()) as ex:
        parts = list(ex.map(work, range(0, n, block)))
    return lml, -0.5 * np.sum(parts, axis=0)                             # -1/2 (b^T dK b - tr(KV^-1 dK))


# --------------------------------------------------------------------------- a22: stochastic Lanczos quadrature
# calculate_random_logdet (gp_lin_alg.py:1103-1181) hands the matrix to imate.logdet(method="slq", lanczos_degree=20,
# min_num_samples=10, max_num_samples=5000, error_rtol=0.01, orthogonalize=0).  imate is an un-vendored, unpinned
# dependency (pyproject.toml:56, `tests` extra) and is not installed here: PARITY UNPINNED against imate itself (its
# probes come from its own RNG and the reference's tests only assert finiteness / rtol 0.1, tests/test_fvgp.py:1897,
# :2282).  What CAN be pinned is the algorithm imate documents (Ubaru, Chen & Saad 2017, "Fast estimation of tr(f(A))
# via stochastic Lanczos quadrature"; imate docs: logdet / slq): Hutchinson with Rademacher probes, m-step Lanczos
# without re-orthogonalisation, Gauss quadrature on the tridiagonal, and the stopping rule
#     z_c * std(samples) / sqrt(ns)  <=  max(error_atol, error_rtol * |mean|),   ns >= min_num_samples,
# with z_c = sqrt(2) erfinv(confidence_level) = 1.96 at imate's default confidence_level = 0.95.
# The probe STREAM below is the product's counter-based hash (splitmix64 of (seed, probe, row)), restated so that both
# sides see identical probes and the comparison is deterministic.
_MASK64 = (1 << 64) - 1


def rademacher_probe(seed, probe, n):
    """+-1 vector of probe number `probe`: bit 0 of splitmix64(seed + golden * (probe * FNV + i + 1))."""
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (np.uint64(probe) * np.uint64(0x100000001B3) + i + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return np.where((z & np.uint64(1)) == 1, 1.0, -1.0)


def slq_samples(KV, degree=20, probes=10, seed=0, probe0=0):
    """One SLQ sample of log det per probe: n * sum_k tau_k^2 log(theta_k) from the degree-step Lanczos tridiagonal."""
    n = KV.shape[0]
    out = np.empty(probes)
    for p in range(probes):
        z = rademacher_probe(seed, probe0 + p, n)
        q_prev, q, beta = np.zeros(n), z / np.sqrt(n), 0.0
        al, be = [], []
        for _ in range(min(degree, n)):
            w = KV @ q - beta * q_prev
            a = float(q @ w)
            w = w - a * q
            beta = float(np.linalg.norm(w))
            al.append(a), be.append(beta)
            if beta < 1e-12 * max(1.0, abs(a)):
                break
            q_prev, q = q, w / beta
        m = len(al)
        T = np.diag(al) + np.diag(be[:m - 1], 1) + np.diag(be[:m - 1], -1)
        lam, vec = np.linalg.eigh(T)
        out[p] = n * np.sum(vec[0, :] ** 2 * np.log(lam))
    return out


SLQ_Z95 = 1.959963984540054          # sqrt(2) * erfinv(0.95)


def slq_converged(samples, error_rtol=0.01, error_atol=0.0, min_num_samples=10):
    """imate's documented stopping rule on the samples drawn so far."""
    ns = len(samples)
    if ns < max(2, min_num_samples):
        return False
    err = SLQ_Z95 * np.std(samples, ddof=1) / np.sqrt(ns)
    return bool(err <= max(error_atol, error_rtol * abs(np.mean(samples))))


def slq_logdet(KV, degree=20, min_num_samples=10, max_num_samples=5000, error_rtol=0.01, seed=0):
    """(estimate, variance of the mean, samples): draw min_num_samples probes, then keep extending the stream by the
    number the current variance says is needed (one step, capped) until the rule holds."""
    samples = slq_samples(KV, degree, min_num_samples, seed)
    while not slq_converged(samples, error_rtol, 0.0, min_num_samples) and len(samples) < max_num_samples:
        need = int(np.ceil((SLQ_Z95 * np.std(samples, ddof=1) / (error_rtol * abs(np.mean(samples)))) ** 2))
        need = min(max_num_samples, max(need, len(samples) + 1))
        samples = np.concatenate([samples, slq_samples(KV, degree, need - len(samples), seed, probe0=len(samples))])
    var = float(np.var(samples, ddof=1) / len(samples)) if len(samples) > 1 else float("nan")
    return float(np.mean(samples)), var, samples
