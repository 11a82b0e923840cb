"""ctypes binding of the C-ABI library `libfvgp_b200.so` (include/fvgp_b200.h).

There is NO CPU fallback: if the library is missing or no CUDA device is present the
compute entry points raise.  torch is used for device memory and streams only.
"""
import ctypes
import os
from ctypes import POINTER, c_double, c_int, c_int32, c_int64, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfvgp_b200.so")

# kernel kinds / fill modes (mirror include/fvgp_b200.h)
K_MATERN32, K_MATERN52, K_SQEXP, K_EXP, K_WENDLAND, K_DISTANCE, K_MATERN52_ROBUST = range(7)
FILL_FULL, FILL_SYMMETRIC, FILL_LOWER = range(3)

_P = c_void_p
_SIGNATURES = {
    "fvgp_version": (c_int, []),
    "fvgp_launch_count": (ctypes.c_ulonglong, []),
    "fvgp_set_bulk_store": (c_int, [c_int]),
    "fvgp_kfill_dense": (c_int, [c_int, c_int, _P, c_int64, _P, c_int64, c_int, c_double, POINTER(c_double),
                                 POINTER(c_double), c_double, _P, _P, c_int64, _P]),
    "fvgp_radial_elementwise": (c_int, [c_int, _P, c_int64, c_double, c_double, _P, _P]),
    "fvgp_kgrad_partials_len": (c_int64, [c_int64, c_int]),
    "fvgp_kgrad_trace_matern32": (c_int, [_P, c_int64, c_int, POINTER(c_double), _P, c_int64, _P, _P,
                                          POINTER(c_double), _P]),
    "fvgp_kgrad_trace_radial": (c_int, [c_int, _P, c_int64, c_int, c_double, POINTER(c_double), c_double, _P, c_int64, _P,
                                        _P, POINTER(c_double), _P]),
    "fvgp_kgrad_block_partials_len": (c_int64, [c_int]),
    "fvgp_kgrad_trace_block_matern32": (c_int, [_P, c_int64, _P, c_int64, c_int, POINTER(c_double), _P, c_int64, _P, _P,
                                                c_int64, _P, _P, _P]),
    "fvgp_kgrad_trace_block_radial": (c_int, [c_int, _P, c_int64, _P, c_int64, c_int, POINTER(c_double), c_double, _P,
                                              c_int64, _P, _P, c_int64, _P, _P, _P]),
    "fvgp_trace_sym_product": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, _P, POINTER(c_double), _P]),
    "fvgp_kgrad_dense_matern32": (c_int, [_P, c_int64, _P, c_int64, c_int, POINTER(c_double), _P, _P]),
    "fvgp_chol_workspace_len": (c_int64, [c_int64]),
    "fvgp_potri_workspace_len": (c_int64, [c_int64]),
    "fvgp_potrf_lower": (c_int, [_P, c_int64, c_int64, _P, _P, _P]),
    "fvgp_potrf_lower_enqueue": (c_int, [_P, c_int64, c_int64, _P, _P, _P]),
    "fvgp_potrs_work_len": (c_int64, [c_int64]),
    "fvgp_potrs_lower": (c_int, [_P, c_int64, c_int64, _P, _P, c_int, c_int64, _P, _P]),
    "fvgp_chol_logdet": (c_int, [_P, c_int64, c_int64, _P, POINTER(c_double), _P]),
    "fvgp_potri_lower": (c_int, [_P, c_int64, c_int64, _P, _P, _P]),
    "fvgp_dgemm_nt": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int, c_int, c_int, c_double, c_double,
                              c_int, _P]),
    "fvgp_dgemm": (c_int, [c_int, c_int, _P, c_int64, _P, c_int64, _P, c_int64, c_int, c_int, c_int, c_double,
                           c_double, c_int, _P]),
    "fvgp_trsm_right_lower_t": (c_int, [_P, c_int64, c_int, _P, c_int64, c_int, _P, _P]),
    "fvgp_trtri_lower": (c_int, [_P, c_int64, c_int64, _P, _P, _P]),
    "fvgp_lauum_lower": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "fvgp_trsv_lower": (c_int, [_P, c_int64, c_int64, _P, _P, c_int, _P, _P]),
    "fvgp_gemv_work_len": (c_int64, [c_int64, c_int64]),
    "fvgp_gemv": (c_int, [c_int, _P, c_int64, c_int, c_int, c_double, _P, _P, _P, _P]),
    "fvgp_dot": (c_int, [_P, _P, c_int64, _P, POINTER(c_double), _P]),
    "fvgp_wendland_aabb_len": (c_int64, [c_int64, c_int]),
    "fvgp_wendland_aabb": (c_int, [_P, c_int64, c_int, _P, _P]),
    "fvgp_wendland_chunk_len": (c_int64, [c_int64, c_int64]),
    "fvgp_wendland_csr_count": (c_int, [_P, c_int64, _P, _P, c_int64, _P, c_int, POINTER(c_double), _P, _P, _P, _P]),
    "fvgp_wendland_csr_fill": (c_int, [_P, c_int64, _P, _P, c_int64, _P, c_int, POINTER(c_double), _P, _P, _P, c_int64,
                                       _P, _P, _P]),
    "fvgp_exclusive_scan_i64": (c_int, [_P, c_int64, _P, _P, POINTER(c_int64), _P]),
    "fvgp_scan_scratch_len": (c_int64, [c_int64]),
    "fvgp_csr_spmv": (c_int, [c_int64, _P, _P, _P, _P, _P, _P]),
    "fvgp_bjacobi_len": (c_int64, [c_int64]),
    "fvgp_bjacobi_build": (c_int, [c_int64, _P, _P, _P, _P, _P]),
    "fvgp_pcg_work_len": (c_int64, [c_int64]),
    "fvgp_pcg": (c_int, [c_int64, _P, _P, _P, _P, _P, _P, c_double, c_int, _P, POINTER(c_int), POINTER(c_double),
                         _P]),
    "fvgp_lanczos_work_len": (c_int64, [c_int64, c_int]),
    "fvgp_lanczos_tridiag": (c_int, [c_int64, _P, _P, _P, c_int, c_int, c_int, c_uint64, _P, POINTER(c_double),
                                     POINTER(c_double), _P]),
    "fvgp_population_slot_len": (c_int64, [c_int64, c_int, c_int]),
    "fvgp_lml_population": (c_int, [c_int, _P, c_int64, c_int, c_int, POINTER(c_double), POINTER(c_double),
                                    POINTER(c_double), POINTER(c_double), _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P,
                                    _P, POINTER(c_double), POINTER(c_double), POINTER(c_double), POINTER(c_int), _P]),
    "fvgp_nccl_attach": (c_int, [ctypes.c_char_p]),
    "fvgp_comm_unique_id": (c_int, [_P]),
    "fvgp_comm_create": (c_int, [_P, c_int, c_int, POINTER(c_void_p)]),
    "fvgp_comm_adopt": (c_int, [_P, c_int, c_int, POINTER(c_void_p)]),
    "fvgp_comm_destroy": (c_int, [_P]),
    "fvgp_comm_allgatherv": (c_int, [_P, _P, POINTER(c_int64), _P]),
    "fvgp_comm_allreduce_sum": (c_int, [_P, _P, c_int64, _P]),
    "fvgp_pcg_sharded_work_len": (c_int64, [c_int64]),
    "fvgp_pcg_sharded": (c_int, [_P, c_int64, POINTER(c_int64), _P, _P, _P, _P, _P, _P, c_double, c_int, _P,
                                 POINTER(c_int), POINTER(c_double), _P]),
    "fvgp_ozaki_available": (c_int, []),
    "fvgp_set_ozaki": (c_int, [c_int]),
    "fvgp_set_ozaki_tri": (c_int, [c_int]),
    "fvgp_ozaki_slices": (c_int, []),
    "fvgp_ozaki_mac_count": (c_uint64, []),
    "fvgp_ozaki_gemm_work_bytes": (c_int64, [c_int, c_int, c_int64, c_int64, c_int64, c_int, c_int64]),
    "fvgp_ozaki_gemm": (c_int, [c_int, c_int, _P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, c_double,
                                c_int, c_int, c_int64, _P, c_int64, _P]),
    "fvgp_set_ozaki_gate": (c_int, [c_int, c_int]),
    "fvgp_ozaki_i8_seconds": (c_double, [c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "fvgp_ozaki_work_bytes": (c_int64, [c_int64, c_int64, c_int64, c_int, c_int64]),
    "fvgp_ozaki_gemm_nt": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, c_double, c_int, c_int64,
                                   c_int, c_int, c_int64, _P, c_int64, _P]),
    "fvgp_bench_fp64_peak": (c_int, [c_int, c_int, c_int, _P, POINTER(c_double), _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed for loading / symbol checks)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise NativeLibraryError("fvgp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def stream_ptr():
    return c_void_p(_torch().cuda.current_stream().cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def dvec(values):
    arr = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
    return arr, arr.ctypes.data_as(POINTER(c_double))


class NonPositiveDefiniteError(Exception):
    """Same contract as fvgp.gp_lin_alg.NonPositiveDefiniteError (gp_lin_alg.py:27-58)."""

    def __init__(self, pivot, n):
        self.pivot = int(pivot)
        super().__init__(
            f"Matrix is not positive definite: the leading minor of order {pivot} (of {n}) is not positive. "
            "Add noise / a nugget or check the kernel and hyperparameters.")


def check(status, what):
    if status < 0:
        raise NativeLibraryError(f"{what} failed with status {status} (see stderr)")
    return status


def dev_empty(shape, dtype=None):
    torch = _torch()
    return torch.empty(shape, dtype=dtype or torch.float64, device="cuda")


def dev_empty_rounded(count, dtype=None):
    """1-d buffer of `count` elements carved from an allocation whose size is rounded up to 1/8 of its power of
    two.  Sizes that drift by a few per cent between evaluations (nnz of the gp2Scale matrix as the length scales
    move) then hit the same block of torch's caching allocator instead of a fresh ~5 ms cudaMalloc each."""
    torch = _torch()
    count = int(count)
    step = max(1 << 18, (1 << max(count, 1).bit_length() - 1) >> 3)
    cap = (count + step - 1) // step * step
    return torch.empty(cap, dtype=dtype or torch.float64, device="cuda")[:count]


def to_dev(a, dtype=None):
    torch = _torch()
    if isinstance(a, torch.Tensor):
        t = a.to(device="cuda", dtype=dtype or torch.float64)
    else:
        t = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype or torch.float64).cuda()
    return t.contiguous()


def padded_ld(n):
    """Leading dimension for an n-column FP64 matrix: even (16-byte rows), 128-byte aligned rows."""
    return int((n + 15) // 16 * 16)


def dev_matrix(rows, cols):
    """rows x cols view on a buffer with padded leading dimension."""
    ld = padded_ld(cols)
    buf = dev_empty((rows, ld))
    return buf, ld
