"""INT8-slice (Ozaki) GEMM on the tcgen05 tensor cores against the DMMA GEMM: accuracy and throughput.
    python tools/ozaki_probe.py [m] [k]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fvgp_b200 import ops  # noqa: E402
from fvgp_b200 import _lib as L  # noqa: E402

lib = L.load()
print("ozaki available:", lib.fvgp_ozaki_available())
rng = np.random.default_rng(0)


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


# ---- accuracy on small problems against float128-free references (FP64 DMMA GEMM and numpy longdouble on a sample)
for m, n, k in ((300, 260, 128), (1024, 1024, 2048)):
    A = rng.standard_normal((m, k)) * np.exp(3 * rng.standard_normal((m, 1)))       # rows of very different scale
    B = rng.standard_normal((n, k)) * np.exp(3 * rng.standard_normal((n, 1)))
    C0 = rng.standard_normal((m, n))
    ref = C0.astype(np.longdouble) - A.astype(np.longdouble) @ B.astype(np.longdouble).T
    scale = np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=1)[None, :] * k
    for S in (6, 7, 8, 9):
        Cd = L.to_dev(C0.copy())
        ops.ozaki_gemm_nt(Cd, L.to_dev(A), L.to_dev(B), sign=-1.0, slices=S, nblock=512)
        err = np.abs(Cd.cpu().numpy() - ref.astype(np.float64))
        print(f"m={m} n={n} k={k} slices={S}: max err / (k rowmax colmax) = {float((err / scale).max()):.2e}, "
              f"max rel err vs |ref| = {float((err / np.maximum(np.abs(ref.astype(np.float64)), 1e-300)).max()):.2e}")
    Cd = L.to_dev(C0.copy())
    ops.dgemm_nt(L.to_dev(A), L.to_dev(B), Cd, alpha=-1.0, beta=1.0)
    err = np.abs(Cd.cpu().numpy() - ref.astype(np.float64))
    print(f"m={m} n={n} k={k} DMMA: max err / (k rowmax colmax) = {float((err / scale).max()):.2e}")
    # SYRK, lower only
    Cd = L.to_dev(C0[:, :m].copy() if n >= m else np.zeros((m, m)))
    if n >= m:
        Ad = L.to_dev(A)
        ops.ozaki_gemm_nt(Cd, Ad, Ad, sign=-1.0, lower=True, slices=8, nblock=256)
        full = C0[:, :m] - A @ A.T
        got = Cd.cpu().numpy()
        low = np.tril(np.ones((m, m), dtype=bool))
        print("  syrk lower: max err", float(np.abs(got - full)[low].max() / np.abs(full).max()),
              "upper untouched:", bool(np.array_equal(got[~low], C0[:, :m][~low])))

# ---- throughput at the shape of the trailing update (m x m x 2048)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
A = torch.randn(m, k, dtype=torch.float64, device="cuda")
C = torch.zeros(m, m, dtype=torch.float64, device="cuda")
t_d = timed(lambda: ops.dgemm_nt(A, A, C, alpha=-1.0, beta=1.0, lower=True))
print(f"SYRK lower m={m} k={k}: DMMA {t_d * 1e3:.1f} ms = {m * m * k / t_d / 1e12:.1f} TFLOP/s (lower half)")
for S in (7, 8, 9):
    for nblock in (4096, 8192):
        t_o = timed(lambda: ops.ozaki_gemm_nt(C, A, A, sign=-1.0, lower=True, slices=S, nblock=nblock), reps=2)
        print(f"  ozaki slices={S} nblock={nblock}: {t_o * 1e3:.1f} ms = {m * m * k / t_o / 1e12:.1f} TFLOP/s FP64-equivalent, "
              f"int8 rate {m * m * k * S * (S + 1) / 2 * 1.0 / t_o / 1e15:.2f} PMAC/s (upper halves of diagonal blocks included)")
