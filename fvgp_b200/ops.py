"""Device operators of the hot path: thin, typed wrappers over the C ABI.

Everything here takes / returns torch CUDA tensors (float64) and enqueues on torch's
current stream.  No arithmetic is done in Python or by torch kernels on these paths.
"""
import ctypes
from ctypes import c_double, c_int, c_int64

import numpy as np

from . import _lib as L

# Optional phase timer (bench.py): CUDA-event pairs on torch's current stream around each operator.
PHASES = None


class _Phase:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PHASES is not None:
            torch = L._torch()
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.nvtx.range_push("fvgp/" + self.name)      # ncu / nsys --nvtx: per-phase ranges inside "timed/"
            self.e0.record()

    def __exit__(self, *exc):
        if PHASES is not None:
            self.e1.record()
            L._torch().cuda.nvtx.range_pop()
            PHASES.setdefault(self.name, []).append((self.e0, self.e1))


def start_phase_timing():
    global PHASES
    PHASES = {}


def stop_phase_timing():
    """Returns {phase: total seconds}; synchronises."""
    global PHASES
    L._torch().cuda.synchronize()
    out = {k: sum(a.elapsed_time(b) for a, b in v) * 1e-3 for k, v in (PHASES or {}).items()}
    PHASES = None
    return out


# ------------------------------------------------------------------------------ dense K
_FOLD = {L.K_MATERN32: np.sqrt(3.0), L.K_MATERN52: np.sqrt(5.0), L.K_SQEXP: np.sqrt(0.5), L.K_EXP: 1.0,
         L.K_WENDLAND: 1.0, L.K_DISTANCE: 1.0, L.K_MATERN52_ROBUST: np.sqrt(5.0)}
CENTRED_LIMIT = 512.0          # bound on |x - centre| * inv_scale * c for the centred fill


def fill_centre(kind, inv_scale, length, bounds):
    """Centre for the whitened K-fill, or None when the scaled half-extent of the points exceeds
    CENTRED_LIMIT (then the fill takes coordinate differences first, like the reference).
    bounds = (lo, hi): per-axis extent of ALL points involved (host arrays)."""
    if bounds is None:
        return None
    lo, hi = (np.asarray(b, dtype=np.float64) for b in bounds)
    if lo.size > 4 or not (np.all(np.isfinite(lo)) and np.all(np.isfinite(hi))):
        return None
    half = 0.5 * (hi - lo) * np.asarray(inv_scale, dtype=np.float64)[:lo.size] * (_FOLD[kind] / float(length))
    if half.size and float(half.max()) > CENTRED_LIMIT:
        return None
    return 0.5 * (lo + hi)


def _require_len(values, need, what, at_least=False):
    """The C side reads `need` doubles through a raw pointer: a short host array must fail here, like the shape
    error the reference's numpy code would raise."""
    have = int(np.size(values))
    if have != need and not (at_least and have > need):
        raise ValueError(f"{what}: expected {'at least ' if at_least else ''}{need} values, got {have}")


def kfill(kind, x1, x2, amp, inv_scale, length=1.0, noise=None, mode=L.FILL_FULL, out=None, bounds=None):
    """K = amp * f(||(x1-x2)*inv_scale|| / length) [+ diag(noise)].  Returns (buffer, ld);
    buffer[:, :n2] is the matrix (gp_prior.py:376-400, kernels.py:16-188, gp_kv.py:640-669).
    bounds: optional per-axis (lo, hi) of the points; enables the centred fast path when safe."""
    lib = L.load()
    n1, n2, dim = x1.shape[0], x2.shape[0], x1.shape[1]
    if out is None:
        buf, ld = L.dev_matrix(n1, n2)
    else:
        buf, ld = out
    _require_len(inv_scale, dim, "inv_scale (one per input dimension; hyperparameters too short?)")
    assert x2.shape[1] == dim, "x1 and x2 must have the same number of columns"
    _, inv_p = L.dvec(inv_scale)
    centre = fill_centre(kind, inv_scale, length, bounds)
    if centre is not None:
        _keep, centre_p = L.dvec(centre)
    else:
        centre_p = None
    with _Phase("kfill"):
        st = lib.fvgp_kfill_dense(kind, mode, L.ptr(x1), n1, L.ptr(x2), n2, dim, float(amp), inv_p, centre_p,
                                  float(length), L.ptr(noise), L.ptr(buf), ld, L.stream_ptr())
    L.check(st, "fvgp_kfill_dense")
    return buf, ld


def kgrad_dense_matern32(x1, x2, theta):
    lib = L.load()
    n1, n2, dim = x1.shape[0], x2.shape[0], x1.shape[1]
    out = L.dev_empty((dim + 1, n1, n2))
    _require_len(theta, dim + 1, "theta (signal variance + one length scale per dimension)", at_least=True)
    _, th = L.dvec(theta)
    L.check(lib.fvgp_kgrad_dense_matern32(L.ptr(x1), n1, L.ptr(x2), n2, dim, th, L.ptr(out), L.stream_ptr()),
            "fvgp_kgrad_dense_matern32")
    return out


def kgrad_trace_matern32(x, theta, kinv_buf, ld, b):
    """sum_ij (Kinv - b b^T)_ij dK_ij/dtheta_h for the default kernel; returns ndarray (dim+1,)."""
    lib = L.load()
    n, dim = x.shape
    partials = L.dev_empty((int(lib.fvgp_kgrad_partials_len(n, dim)),))
    _require_len(theta, dim + 1, "theta (signal variance + one length scale per dimension)", at_least=True)
    _, th = L.dvec(theta)
    out = (c_double * (dim + 1))()
    with _Phase("kgrad_trace"):
        L.check(lib.fvgp_kgrad_trace_matern32(L.ptr(x), n, dim, th, L.ptr(kinv_buf), ld, L.ptr(b), L.ptr(partials),
                                              out, L.stream_ptr()), "fvgp_kgrad_trace_matern32")
    return np.array(out[:], dtype=np.float64)


TRACE_KINDS = (L.K_MATERN32, L.K_MATERN52, L.K_SQEXP, L.K_EXP)


def kgrad_trace_radial(kind, x, amp, inv_scale, length, kinv_buf, ld, b):
    """sum_ij (Kinv - b b^T)_ij dK_ij/dp for p in (amp, inv_scale_1..D, length) of a fused radial kernel;
    returns ndarray (dim + 2,)."""
    lib = L.load()
    n, dim = x.shape
    partials = L.dev_empty((int(lib.fvgp_kgrad_partials_len(n, dim)),))
    _require_len(inv_scale, dim, "inv_scale (one per input dimension)")
    _, inv_p = L.dvec(inv_scale)
    out = (c_double * (dim + 2))()
    with _Phase("kgrad_trace"):
        L.check(lib.fvgp_kgrad_trace_radial(int(kind), L.ptr(x), n, dim, float(amp), inv_p, float(length), L.ptr(kinv_buf),
                                            ld, L.ptr(b), L.ptr(partials), out, L.stream_ptr()), "fvgp_kgrad_trace_radial")
    return np.array(out[:], dtype=np.float64)


def trace_sym_product(kinv_buf, ld, b, dK):
    """sum_ij (Kinv - b b^T)_ij dK_ij for a materialised symmetric dK (device, row-major 2-D tensor)."""
    lib = L.load()
    n = b.numel()
    partials = L.dev_empty((148 * 8 * 2 + 8,))
    out = c_double()
    L.check(lib.fvgp_trace_sym_product(L.ptr(kinv_buf), ld, L.ptr(b), L.ptr(dK), dK.stride(0), n, L.ptr(partials),
                                       ctypes.byref(out), L.stream_ptr()), "fvgp_trace_sym_product")
    return out.value


# ------------------------------------------------------------------------------ dense factorisation
class CholFactor:
    """Lower Cholesky factor resident on the device (+ the tile inverses potrs/potri need)."""

    def __init__(self, buf, ld, n, tileinv):
        self.buf, self.ld, self.n, self.tileinv = buf, ld, n, tileinv
        self.inverted = False

    def lower(self):
        torch = L._torch()
        return torch.tril(self.buf[:, :self.n])


def potrf(buf, ld, n):
    """In-place lower Cholesky (gp_lin_alg.py:237-269).  Raises NonPositiveDefiniteError."""
    lib = L.load()
    torch = L._torch()
    tileinv = L.dev_empty((int(lib.fvgp_chol_workspace_len(n)),))
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    with _Phase("potrf"):
        st = L.check(lib.fvgp_potrf_lower(L.ptr(buf), n, ld, L.ptr(tileinv), L.ptr(info), L.stream_ptr()),
                     "fvgp_potrf_lower")
    if st > 0:
        raise L.NonPositiveDefiniteError(st, n)
    return CholFactor(buf, ld, n, tileinv)


def potrs(factor, rhs_t):
    """Solve KV X = B in place.  rhs_t: (nrhs, n) C-contiguous device tensor = B^T (gp_lin_alg.py:289-328)."""
    lib = L.load()
    assert not factor.inverted
    nrhs, n = rhs_t.shape
    assert n == factor.n and rhs_t.is_contiguous()
    if nrhs > 4 and n % 2:            # GEMM path needs an even row stride
        ldb = n + 1
        padded = L.dev_empty((nrhs, ldb))
        padded[:, :n] = rhs_t
        work = L.dev_empty((int(lib.fvgp_potrs_work_len(n)),))
        L.check(lib.fvgp_potrs_lower(L.ptr(factor.buf), n, factor.ld, L.ptr(factor.tileinv), L.ptr(padded), nrhs, ldb,
                                     L.ptr(work), L.stream_ptr()), "fvgp_potrs_lower")
        rhs_t.copy_(padded[:, :n])
        return rhs_t
    work = L.dev_empty((int(lib.fvgp_potrs_work_len(n)),))
    with _Phase("potrs"):
        L.check(lib.fvgp_potrs_lower(L.ptr(factor.buf), n, factor.ld, L.ptr(factor.tileinv), L.ptr(rhs_t), nrhs, n,
                                     L.ptr(work), L.stream_ptr()), "fvgp_potrs_lower")
    return rhs_t


def chol_logdet(factor):
    lib = L.load()
    scratch = L.dev_empty((1,))
    out = c_double()
    L.check(lib.fvgp_chol_logdet(L.ptr(factor.buf), factor.n, factor.ld, L.ptr(scratch), ctypes.byref(out),
                                 L.stream_ptr()), "fvgp_chol_logdet")
    return out.value


def potri(factor):
    """Lower triangle of the factor buffer <- lower triangle of KV^-1 (gp_lin_alg.py:1558)."""
    lib = L.load()
    work = L.dev_empty((int(lib.fvgp_potri_workspace_len(factor.n)),))
    with _Phase("potri"):
        L.check(lib.fvgp_potri_lower(L.ptr(factor.buf), factor.n, factor.ld, L.ptr(factor.tileinv), L.ptr(work),
                                     L.stream_ptr()), "fvgp_potri_lower")
    factor.inverted = True
    return factor


def population_slots(n, dim, want_grad, batch):
    """Workspace slots for a population of `batch` proposals (= proposals per lock-step chunk, or concurrent streams at
    large N), bounded by 1/4 of the free HBM."""
    lib = L.load()
    torch = L._torch()
    slot_bytes = 8 * int(lib.fvgp_population_slot_len(n, dim, int(want_grad)))
    # n < 6144: lock-step schedule, a slot per proposal of a chunk (more proposals per launch = fewer launches);
    # larger: one stream per slot, each evaluation already fills the GPU
    by_size = 64 if n <= 2048 else (32 if n <= 4096 else (16 if n < 6144 else (4 if n <= 16384 else 2)))
    want = int(max(1, min(batch, by_size)))
    if want * slot_bytes <= (4 << 30):          # small workspaces: no driver query on the hot path (cudaMemGetInfo
        return want                              # costs up to milliseconds and this call sits inside optimiser loops)
    free, _total = torch.cuda.mem_get_info()
    by_mem = max(1, int(0.25 * free) // max(slot_bytes, 1))
    return int(max(1, min(want, by_mem)))


def lml_population(kind, x, amps, inv_scales, lengths, noise, rhs_t, want_grad=False, component=0, bounds=None,
                   slots=None):
    """K-fill -> POTRF -> POTRS -> logdet (-> POTRI -> fused gradient traces) for B proposals of one radial family on
    concurrent streams, ONE host synchronisation (include/fvgp_b200.h: fvgp_lml_population).

    x (n, dim) device; amps (B,), inv_scales (B, dim), lengths (B,) host; noise (n,) device or None; rhs_t (nrhs, n)
    device = (y - m)^T.  Returns alpha (B, nrhs, n), logdet (B,), traces (B, dim + 2) or None, info (B,) int32."""
    lib = L.load()
    torch = L._torch()
    n, dim = x.shape
    amps = np.ascontiguousarray(amps, dtype=np.float64)
    lengths = np.ascontiguousarray(lengths, dtype=np.float64)
    _require_len(inv_scales, len(amps) * dim, "inv_scales (B x dim)")
    _require_len(lengths, len(amps), "lengths (one per proposal)")
    inv_scales = np.ascontiguousarray(inv_scales, dtype=np.float64).reshape(len(amps), dim)
    B = len(amps)
    nrhs = rhs_t.shape[0]
    assert rhs_t.is_contiguous() and rhs_t.shape[1] == n and 1 <= nrhs <= 4
    centre = None
    if bounds is not None:                 # the centred fill must be safe for EVERY proposal
        centre = fill_centre(kind, inv_scales.max(axis=0), float(lengths.min()), bounds)
    if centre is not None:
        _keep, centre_p = L.dvec(centre)
    else:
        centre_p = None
    if slots is None:
        slots = population_slots(n, dim, want_grad, B)
    slots = int(max(1, min(slots, B)))
    work = L.dev_empty((slots * int(lib.fvgp_population_slot_len(n, dim, int(want_grad))),))
    alpha_d = L.dev_empty((B, nrhs, n))
    res_d = L.dev_empty((B, dim + 2))
    info_d = torch.empty(B, dtype=torch.int32, device="cuda")
    alpha = np.empty((B, nrhs, n))
    logdet = np.empty(B)
    traces = np.zeros((B, dim + 2))
    info = np.zeros(B, dtype=np.int32)
    dp = ctypes.POINTER(c_double)
    with _Phase("population"):
        L.check(lib.fvgp_lml_population(int(kind), L.ptr(x), n, dim, B, amps.ctypes.data_as(dp),
                                        inv_scales.ctypes.data_as(dp), lengths.ctypes.data_as(dp), centre_p,
                                        L.ptr(noise), L.ptr(rhs_t), nrhs, int(bool(want_grad)), int(component), slots,
                                        L.ptr(work), L.ptr(alpha_d), L.ptr(res_d), L.ptr(info_d), alpha.ctypes.data_as(dp),
                                        logdet.ctypes.data_as(dp), traces.ctypes.data_as(dp),
                                        info.ctypes.data_as(ctypes.POINTER(c_int)), L.stream_ptr()), "fvgp_lml_population")
    return alpha, logdet, (traces if want_grad else None), info


def dot(a, b):
    lib = L.load()
    scratch = L.dev_empty((1,))
    out = c_double()
    L.check(lib.fvgp_dot(L.ptr(a), L.ptr(b), a.numel(), L.ptr(scratch), ctypes.byref(out), L.stream_ptr()), "fvgp_dot")
    return out.value


def _ld(t):
    """Row stride of a 2-D tensor; torch reports arbitrary strides for size-1 dimensions."""
    return t.stride(0) if t.shape[0] > 1 else t.shape[1] + (t.shape[1] % 2)


def dgemm_nt(A, B, C, alpha=1.0, beta=0.0, lower=False):
    """C = alpha A B^T + beta C on row-major 2-D device tensors (strides taken from the tensors)."""
    lib = L.load()
    m, k = A.shape
    n = B.shape[0]
    assert A.stride(1) == 1 and B.stride(1) == 1 and C.stride(1) == 1
    L.check(lib.fvgp_dgemm_nt(L.ptr(A), _ld(A), L.ptr(B), _ld(B), L.ptr(C), _ld(C), m, n, k,
                              float(alpha), float(beta), int(lower), L.stream_ptr()), "fvgp_dgemm_nt")
    return C


def ozaki_gemm_nt(C, A, B, sign=-1.0, lower=False, diag=0, slices=8, nblock=8192):
    """C += sign * A B^T on the INT8 tensor cores at FP64 accuracy (include/fvgp_b200.h: fvgp_ozaki_gemm_nt).
    C (m, n), A (m, k), B (n, k): row-major 2-D device tensors (strides taken from the tensors); B is A -> SYRK."""
    lib = L.load()
    torch = L._torch()
    m, k = A.shape
    n = B.shape[0]
    same = int(A.data_ptr() == B.data_ptr() and A.shape == B.shape and A.stride(0) == B.stride(0))
    nbytes = int(lib.fvgp_ozaki_work_bytes(m, n, k, int(slices), int(nblock)))
    work = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    L.check(lib.fvgp_ozaki_gemm_nt(L.ptr(C), _ld(C), L.ptr(A), _ld(A), L.ptr(B), _ld(B), m, n, k, float(sign), int(lower),
                                   int(diag), same, int(slices), int(nblock), L.ptr(work), nbytes, L.stream_ptr()),
            "fvgp_ozaki_gemm_nt")
    return C


# ------------------------------------------------------------------------------ gp2Scale sparse
class DeviceCSR:
    """Canonical CSR on the device: int64 indptr, int32 sorted indices, float64 data."""

    def __init__(self, indptr, indices, data, shape):
        self.indptr, self.indices, self.data, self.shape = indptr, indices, data, shape

    @property
    def nnz(self):
        return int(self.data.numel())

    def to_scipy(self):
        import scipy.sparse as sp
        idx_t = np.int32 if (self.nnz < 2 ** 31 and max(self.shape) < 2 ** 31) else np.int64
        return sp.csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy().astype(idx_t, copy=False),
                              self.indptr.cpu().numpy().astype(idx_t)), shape=self.shape)


def wendland_aabb(x):
    lib = L.load()
    n, dim = x.shape
    boxes = L.dev_empty((max(1, int(lib.fvgp_wendland_aabb_len(n, dim))),))
    L.check(lib.fvgp_wendland_aabb(L.ptr(x), n, dim, L.ptr(boxes), L.stream_ptr()), "fvgp_wendland_aabb")
    return boxes


def wendland_csr(x1, x2, theta, noise=None, boxes1=None, boxes2=None, stats=None):
    """k(x1, x2) for the compact-support Wendland kernel as canonical CSR (bit-exact pattern).

    Replaces the per-block dask tasks + host assembly (gp2Scale_covariance.py:136-287); `noise`
    fuses K + diag(V) (gp_kv.py:655-661).  stats: optional int64 device tensor (2,), receives += the number
    of 32x32 tile pairs that survived the box culls and the number of candidate point pairs tested."""
    lib = L.load()
    torch = L._torch()
    n1, dim = x1.shape
    n2 = x2.shape[0]
    same = x1.data_ptr() == x2.data_ptr() and n1 == n2
    if boxes1 is None:
        boxes1 = wendland_aabb(x1)
    if boxes2 is None:
        boxes2 = boxes1 if same else wendland_aabb(x2)
    _, th = L.dvec(theta)
    counts = torch.zeros(max(n1, 1), dtype=torch.int64, device="cuda")
    st = L.stream_ptr()
    chunk = torch.empty(int(lib.fvgp_wendland_chunk_len(n1, n2)), dtype=torch.int32, device="cuda")
    L.check(lib.fvgp_wendland_csr_count(L.ptr(x1), n1, L.ptr(boxes1), L.ptr(x2), n2, L.ptr(boxes2), dim, th,
                                        L.ptr(counts), L.ptr(chunk), L.ptr(stats), st), "fvgp_wendland_csr_count")
    indptr = torch.empty(n1 + 1, dtype=torch.int64, device="cuda")
    scratch = torch.empty(int(lib.fvgp_scan_scratch_len(n1)), dtype=torch.int64, device="cuda")
    total = c_int64()
    L.check(lib.fvgp_exclusive_scan_i64(L.ptr(counts), n1, L.ptr(indptr), L.ptr(scratch), ctypes.byref(total), st),
            "fvgp_exclusive_scan_i64")
    nnz = total.value
    indices = L.dev_empty_rounded(nnz, torch.int32)
    data = L.dev_empty_rounded(nnz, torch.float64)
    if nnz:
        L.check(lib.fvgp_wendland_csr_fill(L.ptr(x1), n1, L.ptr(boxes1), L.ptr(x2), n2, L.ptr(boxes2), dim, th,
                                           L.ptr(indptr), L.ptr(chunk), L.ptr(noise), 0, L.ptr(indices), L.ptr(data), st),
                "fvgp_wendland_csr_fill")
    return DeviceCSR(indptr, indices, data, (n1, n2))


def spmv(A, x, y=None):
    lib = L.load()
    if y is None:
        y = L.dev_empty((A.shape[0],))
    L.check(lib.fvgp_csr_spmv(A.shape[0], L.ptr(A.indptr), L.ptr(A.indices), L.ptr(A.data), L.ptr(x), L.ptr(y),
                              L.stream_ptr()), "fvgp_csr_spmv")
    return y


def bjacobi(A):
    lib = L.load()
    blocks = L.dev_empty((int(lib.fvgp_bjacobi_len(A.shape[0])),))
    L.check(lib.fvgp_bjacobi_build(A.shape[0], L.ptr(A.indptr), L.ptr(A.indices), L.ptr(A.data), L.ptr(blocks),
                                   L.stream_ptr()), "fvgp_bjacobi_build")
    return blocks


def pcg(A, b, x0=None, rtol=1e-5, maxiter=None, precond=None):
    """scipy.sparse.linalg.cg semantics (gp_lin_alg.py:1284-1288). Returns (x, info, iters, relres)."""
    lib = L.load()
    torch = L._torch()
    n = A.shape[0]
    x = torch.zeros(n, dtype=torch.float64, device="cuda") if x0 is None else x0.clone().contiguous()
    work = L.dev_empty((int(lib.fvgp_pcg_work_len(n)),))
    iters, relres = c_int(), c_double()
    if maxiter is None:
        maxiter = 10 * n
    st = L.check(lib.fvgp_pcg(n, L.ptr(A.indptr), L.ptr(A.indices), L.ptr(A.data), L.ptr(precond), L.ptr(b), L.ptr(x),
                              float(rtol), int(maxiter), L.ptr(work), ctypes.byref(iters), ctypes.byref(relres),
                              L.stream_ptr()), "fvgp_pcg")
    return x, st, iters.value, relres.value


def slq_logdet(A, degree=20, probes=30, seed=0, probe0=0):
    """Stochastic Lanczos quadrature estimate of log det A (gp_lin_alg.py:1103-1181, imate slq).

    Device: Lanczos three-term recurrences (SpMV bound).  Host: eigen-decomposition of the
    `degree` x `degree` tridiagonals (microseconds).  Probes probe0 .. probe0+probes-1 of the counter-based
    Rademacher stream advance in lock step, up to 16 per sweep over the matrix (SpMM).
    Returns (estimate, variance_of_mean, samples)."""
    lib = L.load()
    n = A.shape[0]
    degree = int(min(degree, n))
    work = L.dev_empty((int(lib.fvgp_lanczos_work_len(n, degree)),))
    alpha = np.zeros(probes * degree)
    beta = np.zeros(probes * degree)
    L.check(lib.fvgp_lanczos_tridiag(n, L.ptr(A.indptr), L.ptr(A.indices), L.ptr(A.data), degree, int(probe0), probes,
                                     int(seed), L.ptr(work), alpha.ctypes.data_as(ctypes.POINTER(c_double)),
                                     beta.ctypes.data_as(ctypes.POINTER(c_double)), L.stream_ptr()),
            "fvgp_lanczos_tridiag")
    samples = np.empty(probes)
    for p in range(probes):
        a = alpha[p * degree:(p + 1) * degree]
        b = beta[p * degree:(p + 1) * degree - 1]
        m = degree
        small = np.nonzero(b < 1e-12 * max(1.0, np.abs(a).max()))[0]      # invariant subspace found early
        if small.size:
            m = int(small[0]) + 1
        T = np.diag(a[:m]) + np.diag(b[:m - 1], 1) + np.diag(b[:m - 1], -1)
        lam, vec = np.linalg.eigh(T)
        samples[p] = n * np.sum(vec[0, :] ** 2 * np.log(np.maximum(lam, 1e-300)))
    est = float(samples.mean())
    var = float(samples.var(ddof=1) / probes) if probes > 1 else float("nan")
    return est, var, samples
