"""TEST INFRASTRUCTURE ONLY: a torch-CPU stand-in for fvgp_b200.sharded.CudaLocalOps.

It lets the gloo tests execute the block-cyclic choreography of fvgp_b200/sharded.py (which block goes
where, which collective carries it, which local product updates it) without a GPU.  Each method restates
the SEMANTICS of one C-ABI entry point with torch.linalg on CPU tensors; the product never imports this
module, and the GPU tests run the same choreography on the real kernels."""
import numpy as np
import torch

from oracle import fvgp_oracle as orc

K_MATERN32 = 0


class CpuLocalOps:
    device = "cpu"

    def empty(self, *shape):
        return torch.full(shape, float("nan"), dtype=torch.float64)      # catch reads of unwritten memory

    def zeros(self, *shape):
        return torch.zeros(shape, dtype=torch.float64)

    def upload(self, a):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)

    def tileinv_len(self, n):
        return 8

    def fill(self, kind, x1, x2, amp, inv_scale, length, noise, out, centre):
        if kind == K_MATERN32 and length == 1.0:
            hps = np.concatenate([[amp], 1.0 / np.asarray(inv_scale)])
            k = orc.default_kernel(x1.numpy(), x2.numpy(), hps)
        else:                                                             # the other radial families, oracle formulas
            d = orc.anisotropic_distance_matrix(x1.numpy(), x2.numpy(), 1.0 / np.asarray(inv_scale))
            name = {0: "matern32", 1: "matern52", 2: "se", 3: "exp"}[int(kind)]
            k = amp * orc.RADIAL[name](d, length)
        if noise is not None:
            k[np.arange(len(noise)), np.arange(len(noise))] += noise.numpy()
        out[:x1.shape[0], :x2.shape[0]] = torch.from_numpy(k)

    def potrf(self, A, n, tileinv):
        a = torch.tril(A[:n, :n])
        a = a + torch.tril(a, -1).T
        try:
            Lf = torch.linalg.cholesky(a)
        except Exception:
            return 1
        A[:n, :n] = Lf                                                    # explicit zeros above the diagonal
        return 0

    def trsm_rlt(self, B, m, Lf, n, tileinv):
        Ltri = torch.tril(Lf[:n, :n])
        B[:m, :n] = torch.linalg.solve_triangular(Ltri, B[:m, :n].T.contiguous(), upper=False).T

    def gemm(self, a_mn, b_mn, A, B, C, m, n, k, alpha, beta, flags=0):
        a = A[:k, :m].T if a_mn else A[:m, :k]
        b = B[:k, :n] if b_mn else B[:n, :k].T
        prod = a @ b
        C[:m, :n] = alpha * prod + (beta * C[:m, :n] if beta != 0.0 else 0.0)

    def trtri(self, A, n, tileinv):
        A[:n, :n] = torch.linalg.inv(torch.tril(A[:n, :n]))

    def lauum(self, A, n):
        m = torch.tril(A[:n, :n])
        A[:n, :n] = torch.tril(m.T @ m)

    def trsv(self, Lf, n, tileinv, b, transpose):
        Ltri = torch.tril(Lf[:n, :n])
        if transpose:
            b[:n] = torch.linalg.solve_triangular(Ltri.T, b[:n, None], upper=True)[:, 0]
        else:
            b[:n] = torch.linalg.solve_triangular(Ltri, b[:n, None], upper=False)[:, 0]

    def gemv(self, transpose, A, m, n, alpha, x, y):
        if transpose:
            y[:n] += alpha * (A[:m, :n].T @ x[:m])
        else:
            y[:m] += alpha * (A[:m, :n] @ x[:n])

    def logdet(self, Lf, n):
        return float(2.0 * torch.log(torch.abs(torch.diagonal(Lf[:n, :n]))).sum())

    def trace_block(self, x1, x2, theta, W, m, n, b1, b2, diag_rows, accum):
        theta = np.asarray(theta)
        dK = orc.default_kernel_gradient(x1.numpy()[:m], x2.numpy()[:n], theta)            # (H, m, n)
        w = W[:m, :n].numpy() - np.outer(b1.numpy()[:m], b2.numpy()[:n])
        weight = np.full((m, n), 2.0)
        if diag_rows:
            r, c = np.arange(diag_rows)[:, None], np.arange(n)[None, :]
            weight[:diag_rows] = np.where(c > r, 0.0, np.where(c == r, 1.0, 2.0))
        w = np.where(weight == 0.0, 0.0, w) * weight
        accum += torch.from_numpy(np.einsum("hij,ij->h", dK, w))

    def trace_block_radial(self, kind, x1, x2, inv_scale, length, W, m, n, b1, b2, diag_rows, accum):
        """Raw sums R_0 = sum W f(u), R_i = sum W h(u) q_i of the radial families (fvgp_kgrad_trace_block_radial):
        q_i = squared difference along axis i in coordinates scaled by inv_scale_i * c_KIND / length."""
        fold = {0: np.sqrt(3.0), 1: np.sqrt(5.0), 2: np.sqrt(0.5), 3: 1.0}[int(kind)] / float(length)
        a1 = x1.numpy()[:m] * (np.asarray(inv_scale) * fold)
        a2 = x2.numpy()[:n] * (np.asarray(inv_scale) * fold)
        q = (a1[:, None, :] - a2[None, :, :]) ** 2                                           # (m, n, D)
        s = q.sum(-1)
        a = np.sqrt(np.maximum(s, 1e-300))
        e = np.exp(-a)
        if kind == 2:
            f, h = np.exp(-s), 2.0 * np.exp(-s)
        elif kind == 0:
            f, h = (1 + a) * e, e
        elif kind == 1:
            f, h = (1 + a + s / 3.0) * e, (1 + a) * e / 3.0
        else:
            f, h = e, e / a
        w = W[:m, :n].numpy() - np.outer(b1.numpy()[:m], b2.numpy()[:n])
        weight = np.full((m, n), 2.0)
        if diag_rows:
            r, c = np.arange(diag_rows)[:, None], np.arange(n)[None, :]
            weight[:diag_rows] = np.where(c > r, 0.0, np.where(c == r, 1.0, 2.0))
        w = np.where(weight == 0.0, 0.0, w) * weight
        raw = np.concatenate([[np.sum(w * f)], np.einsum("ij,ijd->d", w * h, q)])
        accum[:len(raw)] += torch.from_numpy(raw)
