"""Numerical prototype (numpy, host) for DESIGN.md section 8 item 1: trailing updates of the blocked Cholesky through an
Ozaki-type error-free split into 6-bit integer slices (the INT8 tcgen05 path on B200), instead of FP64 DMMA.

NOT product code and not on any product path: it answers one design question before any kernel is written -- how many
slices does the C2 workload (ARD Matern-3/2 + 1e-2 noise) need for the LML to stay within the 1e-8 parity budget?

    A (m x k panel of L)  ->  row scaling 2^e_i,  A / 2^e_i = sum_t S_t 2^(-6 t),  S_t integer, |S_t| <= 64
    A A^T = sum_{t,u} 2^(-6 (t + u)) 2^(e_i + e_j) S_t S_u^T      (integer products: exact in INT32 for k < 2^19)
    pairs with t + u > p + 1 are dropped  ->  p (p + 1) / 2 integer GEMMs per FP64 GEMM

Integer-valued float64 arrays stand in for the INT8 operands (their products are exact in float64 at these sizes)."""
import sys
import time

import numpy as np
import scipy.linalg as sla


def slices(A, p):
    """Row-scaled 6-bit slices of A: returns (list of integer-valued arrays, row exponents)."""
    amax = np.max(np.abs(A), axis=1)
    e = np.ceil(np.log2(np.maximum(amax, 1e-300))) + 1          # |A / 2^e| < 1/2 ... safe headroom
    R = A / np.exp2(e)[:, None]
    out = []
    for _ in range(p):
        R = R * 64.0
        S = np.trunc(R)
        out.append(S)
        R = R - S
    return out, e


def syrk_sliced(A, p):
    """A A^T through p slices (pairs with t + u <= p + 1, 1-based)."""
    S, e = slices(A, p)
    acc = np.zeros((A.shape[0], A.shape[0]))
    products = 0
    for g in range(2, p + 2):                                    # g = t + u
        G = np.zeros_like(acc)
        for t in range(1, g):
            u = g - t
            if t <= p and u <= p:
                G += S[t - 1] @ S[u - 1].T                       # exact integer GEMM (INT8 x INT8 -> INT32)
                products += 1
        acc += G * 2.0 ** (-6 * g)                               # one FP64 conversion + scale per group
    return acc * np.exp2(e)[:, None] * np.exp2(e)[None, :], products


def blocked_cholesky(K, nb, p=None):
    """Right-looking blocked Cholesky; the trailing update in FP64 (p=None) or through p integer slices."""
    A = np.array(K, copy=True)
    n = len(A)
    for j in range(0, n, nb):
        w = min(nb, n - j)
        A[j:j + w, j:j + w] = np.linalg.cholesky(A[j:j + w, j:j + w])
        if j + w < n:
            A[j + w:, j:j + w] = sla.solve_triangular(A[j:j + w, j:j + w], A[j + w:, j:j + w].T, lower=True).T
            panel = A[j + w:, j:j + w]
            upd = panel @ panel.T if p is None else syrk_sliced(panel, p)[0]
            A[j + w:, j + w:] -= upd
    return np.tril(A)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    rng = np.random.default_rng(2)
    x = rng.random((n, 3))
    y = np.sin(5 * x[:, 0]) * np.cos(3 * x[:, 1]) + x[:, 2] + 0.1 * rng.standard_normal(n)
    th = np.array([1.0, .3, .4, .5])
    d = np.sqrt(sum(((x[:, i][:, None] - x[:, i][None, :]) / th[1 + i]) ** 2 for i in range(3)))
    K = th[0] * (1 + np.sqrt(3) * d) * np.exp(-np.sqrt(3) * d) + 1e-2 * np.eye(n)
    ym = y - y.mean()

    def lml(L):
        a = sla.cho_solve((L, True), ym)
        return -0.5 * (ym @ a + 2 * np.sum(np.log(np.diag(L))) + n * np.log(2 * np.pi))
    ref = lml(np.linalg.cholesky(K))
    print(f"N = {n}, cond(KV) = {np.linalg.cond(K):.2e}, LML = {ref:.10f}")
    print(f"blocked FP64 (nb 256): rel LML error {abs(lml(blocked_cholesky(K, 256)) / ref - 1):.2e}")
    A = rng.standard_normal((512, 256))
    for p in (4, 5, 6, 7, 8, 9):
        t0 = time.time()
        L = blocked_cholesky(K, 256, p)
        err = abs(lml(L) / ref - 1)
        prod, cnt = syrk_sliced(A, p)
        gerr = np.max(np.abs(prod - A @ A.T)) / np.max(np.abs(A @ A.T))
        print(f"p = {p} slices ({cnt:2d} integer GEMMs per FP64 GEMM): rel LML error {err:.2e}, "
              f"GEMM error / max|C| {gerr:.1e}, INT8-equivalent rate at 3.3 POPS: {3300 / cnt:.0f} TFLOP/s "
              f"(DMMA: 34)   [{time.time() - t0:.0f} s]")


if __name__ == "__main__":
    main()
