"""The binding a reference maintainer would add: ctypes stubs over `libfvgp_b200.so` for the operator seams the
UNMODIFIED reference already exposes (INTEGRATION.md sections 1-3).  Nothing here imports the `fvgp_b200` Python
package -- only the C ABI of include/fvgp_b200.h -- and nothing in the reference is patched:

    import fvgp                                   # the reference
    from fvgp_reference_plugin import b200_linalg_mode, b200_default_kernel, b200_wendland_gp2Scale
    gp = fvgp.GP(x, y, init_hyperparameters=h, noise_variances=v,
                 kernel_function=b200_default_kernel,          # kernel seam      gp_prior.py:61, 217-224
                 linalg_mode=b200_linalg_mode)                 # linalg seam      gp.py:274-281, gp_kv.py:457-715

Executed against the reference installed under baseline/_ref by tests/test_gpu_reference_plugin.py."""
import ctypes
import os

import numpy as np
import scipy.sparse
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("FVGP_B200_LIB", os.path.join(_HERE, "..", "fvgp_b200", "lib", "libfvgp_b200.so"))
_b200 = ctypes.CDLL(LIB)
_P, _i64, _dbl, _int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int
_DP = ctypes.POINTER(_dbl)
_b200.fvgp_potrf_lower.argtypes = [_P, _i64, _i64, _P, _P, _P]
_b200.fvgp_potrs_lower.argtypes = [_P, _i64, _i64, _P, _P, _int, _i64, _P, _P]
_b200.fvgp_chol_logdet.argtypes = [_P, _i64, _i64, _P, _DP, _P]
_b200.fvgp_chol_workspace_len.restype = _i64
_b200.fvgp_chol_workspace_len.argtypes = [_i64]
_b200.fvgp_potrs_work_len.restype = _i64
_b200.fvgp_potrs_work_len.argtypes = [_i64]
_b200.fvgp_kfill_dense.argtypes = [_int, _int, _P, _i64, _P, _i64, _int, _dbl, _DP, _DP, _dbl, _P, _P, _i64, _P]
_b200.fvgp_wendland_aabb_len.restype = _i64
_b200.fvgp_wendland_aabb_len.argtypes = [_i64, _int]
_b200.fvgp_wendland_aabb.argtypes = [_P, _i64, _int, _P, _P]
_b200.fvgp_wendland_chunk_len.restype = _i64
_b200.fvgp_wendland_chunk_len.argtypes = [_i64, _i64]
_b200.fvgp_wendland_csr_count.argtypes = [_P, _i64, _P, _P, _i64, _P, _int, _DP, _P, _P, _P, _P]
_b200.fvgp_wendland_csr_fill.argtypes = [_P, _i64, _P, _P, _i64, _P, _int, _DP, _P, _P, _P, _i64, _P, _P, _P]
_b200.fvgp_exclusive_scan_i64.argtypes = [_P, _i64, _P, _P, ctypes.POINTER(_i64), _P]
_b200.fvgp_scan_scratch_len.restype = _i64
_b200.fvgp_scan_scratch_len.argtypes = [_i64]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class NonPositiveDefiniteError(Exception):
    """Same contract as fvgp.gp_lin_alg.NonPositiveDefiniteError (gp_lin_alg.py:27-58)."""


# ---- linear-algebra seam: linalg_mode = (f_factor, f_solve, f_logdet) --------------------------------------------
class B200Factor:
    def __init__(self, KV):                # KV: ndarray (N, N) from the reference's addKV (gp_kv.py:640-669)
        KV = np.asarray(KV, dtype=np.float64)
        n = KV.shape[0]
        ld = (n + 15) // 16 * 16
        self.n, self.ld = n, ld
        self.buf = torch.empty((n, ld), dtype=torch.float64, device="cuda")
        self.buf[:, :n] = torch.as_tensor(KV, device="cuda")
        self.tileinv = torch.empty(_b200.fvgp_chol_workspace_len(n), dtype=torch.float64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        st = _b200.fvgp_potrf_lower(self.buf.data_ptr(), n, ld, self.tileinv.data_ptr(), info.data_ptr(), _stream())
        if st > 0:
            raise NonPositiveDefiniteError(f"leading minor of order {st} (of {n}) is not positive")
        if st < 0:
            raise RuntimeError("fvgp_potrf_lower failed")


def b200_factor(KV):                       # replaces calculate_Chol_factor, gp_lin_alg.py:237-269
    return B200Factor(KV)


def b200_solve(f, b):                      # replaces calculate_Chol_solve, gp_lin_alg.py:289-328
    b = np.asarray(b, dtype=np.float64)
    rhs = torch.as_tensor(np.ascontiguousarray(b.reshape(f.n, -1).T), device="cuda")     # (r, N): one RHS per row
    ldb = f.n
    if rhs.shape[0] > 4 and f.n % 2:       # the many-right-hand-side (GEMM) path needs an even row stride
        ldb = f.n + 1
        padded = torch.zeros((rhs.shape[0], ldb), dtype=torch.float64, device="cuda")
        padded[:, :f.n] = rhs
        rhs = padded
    work = torch.empty(_b200.fvgp_potrs_work_len(f.n), dtype=torch.float64, device="cuda")
    st = _b200.fvgp_potrs_lower(f.buf.data_ptr(), f.n, f.ld, f.tileinv.data_ptr(), rhs.data_ptr(), rhs.shape[0], ldb,
                                work.data_ptr(), _stream())
    if st != 0:
        raise RuntimeError("fvgp_potrs_lower failed")
    return rhs[:, :f.n].cpu().numpy().T.reshape(b.shape)


def b200_logdet(f):                        # replaces calculate_Chol_logdet, gp_lin_alg.py:331-360
    out, scratch = _dbl(), torch.empty(1, dtype=torch.float64, device="cuda")
    _b200.fvgp_chol_logdet(f.buf.data_ptr(), f.n, f.ld, scratch.data_ptr(), ctypes.byref(out), _stream())
    return out.value


b200_linalg_mode = (b200_factor, b200_solve, b200_logdet)


# ---- kernel seam: kernel_function(x1, x2, hps) ---------------------------------------------------------------------
def b200_default_kernel(x1, x2, hps):      # replaces GPprior._default_kernel, gp_prior.py:376-400
    x1, x2 = np.ascontiguousarray(x1, dtype=np.float64), np.ascontiguousarray(x2, dtype=np.float64)
    n1, n2, d = len(x1), len(x2), x1.shape[1]
    ld = (n2 + 15) // 16 * 16
    K = torch.empty((n1, ld), dtype=torch.float64, device="cuda")
    inv = (_dbl * d)(*(1.0 / np.asarray(hps, dtype=np.float64)[1:1 + d]))
    a, b = torch.as_tensor(x1, device="cuda"), torch.as_tensor(x2, device="cuda")
    # kind 0 = Matern-3/2, mode 0 = full; h_centre = NULL keeps the reference's operation order ((x1-x2)/l per entry)
    st = _b200.fvgp_kfill_dense(0, 0, a.data_ptr(), n1, b.data_ptr(), n2, d, float(hps[0]), inv, None, 1.0, None,
                                K.data_ptr(), ld, _stream())
    if st != 0:
        raise RuntimeError("fvgp_kfill_dense failed")
    return K[:, :n2].cpu().numpy()


def b200_wendland_gp2Scale(x1, x2, hps):   # replaces wendland_anisotropic_gp2Scale_cpu, kernels.py:502-528
    """Returns scipy.sparse.csr_matrix; block_to_coo (gp2Scale_covariance.py:136-148) passes sparse blocks through."""
    x1, x2 = np.ascontiguousarray(x1, dtype=np.float64), np.ascontiguousarray(x2, dtype=np.float64)
    n1, n2, d = len(x1), len(x2), x1.shape[1]
    a, b = torch.as_tensor(x1, device="cuda"), torch.as_tensor(x2, device="cuda")
    th = (_dbl * (d + 1))(*np.asarray(hps, dtype=np.float64)[:d + 1])

    def boxes(t, n):
        out = torch.empty(max(1, _b200.fvgp_wendland_aabb_len(n, d)), dtype=torch.float64, device="cuda")
        _b200.fvgp_wendland_aabb(t.data_ptr(), n, d, out.data_ptr(), _stream())
        return out
    ba, bb = boxes(a, n1), boxes(b, n2)
    counts = torch.zeros(max(n1, 1), dtype=torch.int64, device="cuda")
    chunk = torch.empty(_b200.fvgp_wendland_chunk_len(n1, n2), dtype=torch.int32, device="cuda")
    _b200.fvgp_wendland_csr_count(a.data_ptr(), n1, ba.data_ptr(), b.data_ptr(), n2, bb.data_ptr(), d, th,
                                  counts.data_ptr(), chunk.data_ptr(), None, _stream())
    indptr = torch.empty(n1 + 1, dtype=torch.int64, device="cuda")
    scratch = torch.empty(_b200.fvgp_scan_scratch_len(n1), dtype=torch.int64, device="cuda")
    nnz = _i64()
    _b200.fvgp_exclusive_scan_i64(counts.data_ptr(), n1, indptr.data_ptr(), scratch.data_ptr(), ctypes.byref(nnz), _stream())
    idx = torch.empty(max(nnz.value, 1), dtype=torch.int32, device="cuda")
    val = torch.empty(max(nnz.value, 1), dtype=torch.float64, device="cuda")
    if nnz.value:
        _b200.fvgp_wendland_csr_fill(a.data_ptr(), n1, ba.data_ptr(), b.data_ptr(), n2, bb.data_ptr(), d, th,
                                     indptr.data_ptr(), chunk.data_ptr(), None, 0, idx.data_ptr(), val.data_ptr(), _stream())
    return scipy.sparse.csr_matrix((val[:nnz.value].cpu().numpy(), idx[:nnz.value].cpu().numpy(),
                                    indptr.cpu().numpy().astype(np.int32)), shape=(n1, n2))
