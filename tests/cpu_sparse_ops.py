"""torch-CPU / numpy stand-in for fvgp_b200.sharded_sparse.CudaSparseOps -- TEST INFRASTRUCTURE ONLY.

Lets the choreography of the multi-GPU gp2Scale path (slab edges, count balance, in-place all-gather of the strips,
row-sharded CG contract, probe split) run under gloo on CPU.  The numerical pieces come from the oracle
(oracle/fvgp_oracle.py); nothing here is reachable from the product."""
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

from oracle import fvgp_oracle as orc


class CpuCSR:
    def __init__(self, indptr, indices, data, n):
        self.indptr, self.indices, self.data, self.shape = indptr, indices, data, (n, n)

    @property
    def nnz(self):
        return int(self.data.numel())

    def to_scipy(self):
        return sp.csr_matrix((self.data.numpy(), self.indices.numpy(), self.indptr.numpy().astype(np.int64)),
                             shape=self.shape)


class CpuSparseOps:
    device = "cpu"

    def __init__(self):
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.count_calls = 0

    def aabb(self, x):
        return None

    def zeros_i64(self, n):
        return torch.zeros(max(int(n), 1), dtype=torch.int64)

    def _block(self, x, row0, nrows, theta):
        xs = x.numpy()
        blk = orc.wendland_block(xs[row0:row0 + nrows], xs, np.asarray(theta, dtype=float))
        return blk

    def count_slab(self, x, boxes, row0, nrows, theta, counts):
        self.count_calls += 1
        blk = self._block(x, row0, nrows, theta)
        counts[row0:row0 + nrows] = torch.as_tensor(np.count_nonzero(blk, axis=1), dtype=torch.int64)
        return (row0, nrows, blk)                                  # "chunk": what the fill pass needs from the count pass

    def scan(self, counts, n):
        indptr = torch.zeros(n + 1, dtype=torch.int64)
        indptr[1:] = torch.cumsum(counts[:n], 0)
        return indptr, int(indptr[-1])

    def alloc_csr(self, nnz):
        return torch.full((nnz,), -1, dtype=torch.int32), torch.full((nnz,), float("nan"), dtype=torch.float64)

    def fill_slab(self, x, boxes, row0, nrows, theta, indptr, chunk, noise, indices, data):
        c_row0, c_nrows, blk = chunk
        assert (c_row0, c_nrows) == (row0, nrows), "the fill pass must use the count pass of the SAME slab"
        r, c = np.nonzero(blk)
        v = blk[r, c].copy()
        if noise is not None:
            diag = c == r + row0
            v[diag] += noise.numpy()[r[diag] + row0]
        a, b = int(indptr[row0]), int(indptr[row0 + nrows])
        assert b - a == len(v)
        indices[a:b] = torch.as_tensor(c.astype(np.int32))
        data[a:b] = torch.as_tensor(v)

    def allgatherv(self, t, elem_offsets):
        if self.world == 1:
            return
        for r in range(self.world):
            a, b = int(elem_offsets[r]), int(elem_offsets[r + 1])
            if b > a:
                part = t[a:b].clone()
                dist.broadcast(part, r)
                t[a:b] = part

    def prefix_at(self, indptr, rows):
        return [int(indptr[r]) for r in rows]

    def inclusive_cumsum_host_view(self, counts, n):
        return np.cumsum(counts[:n].numpy())

    def bjacobi(self, csr):
        return None

    def pcg(self, rows, csr, precond, b, x0, rtol, maxiter):
        """Row-sharded CG with the contract of fvgp_pcg_sharded (scipy stopping rule), collectives through gloo."""
        n = csr.shape[0]
        r0, r1 = int(rows[self.rank]), int(rows[self.rank + 1])
        A = csr.to_scipy()[r0:r1]
        maxiter = 10 * n if maxiter is None else maxiter

        def allsum(v):
            t = torch.as_tensor(np.atleast_1d(np.asarray(v, dtype=float)))
            if self.world > 1:
                dist.all_reduce(t)
            return t.numpy()
        x = np.zeros(n) if x0 is None else x0.numpy().copy()
        bb = b.numpy()
        r = bb[r0:r1] - A @ x
        bn2, = allsum(bb[r0:r1] @ bb[r0:r1])
        atol = rtol * np.sqrt(bn2)
        p = np.zeros(n)
        rho_old, iters, info = 0.0, 0, 1
        while True:
            rr, rz = allsum([r @ r, r @ r])
            if np.sqrt(rr) < atol:
                info = 0
                break
            if iters >= maxiter:
                break
            beta = rz / rho_old if iters > 0 else 0.0
            rho_old = rz
            p[r0:r1] = r + beta * p[r0:r1]
            pt = torch.as_tensor(p)
            self.allgatherv(pt, rows)
            p = pt.numpy()
            q = A @ p
            pq, = allsum(p[r0:r1] @ q)
            alpha = rz / pq
            x[r0:r1] += alpha * p[r0:r1]
            r -= alpha * q
            iters += 1
        xt = torch.as_tensor(x)
        self.allgatherv(xt, rows)
        return xt, info, iters, float(np.sqrt(rr / bn2)) if bn2 > 0 else 0.0

    def slq_samples(self, csr, degree, probe0, count, seed):
        A = csr.to_scipy()
        n = A.shape[0]
        out = np.empty(count)
        for k in range(count):
            z = np.random.default_rng([seed, probe0 + k]).choice([-1.0, 1.0], size=n)      # probe stream indexed by probe
            q_prev, q = np.zeros(n), z / np.sqrt(n)
            al, be = [], []
            beta = 0.0
            for j in range(min(degree, n)):
                w = A @ q - beta * q_prev
                a = q @ w
                w -= a * q
                beta = np.linalg.norm(w)
                al.append(a)
                be.append(beta)
                if beta < 1e-12:
                    break
                q_prev, q = q, w / beta
            m = len(al)
            T = np.diag(al) + np.diag(be[:m - 1], 1) + np.diag(be[:m - 1], -1)
            lam, vec = np.linalg.eigh(T)
            out[k] = n * np.sum(vec[0] ** 2 * np.log(lam))
        return out

    def gather_samples(self, mine, offsets):
        if self.world == 1:
            return mine
        buf = torch.zeros(int(offsets[-1]), dtype=torch.float64)
        a, b = int(offsets[self.rank]), int(offsets[self.rank + 1])
        if b > a:
            buf[a:b] = torch.as_tensor(mine)
        dist.all_reduce(buf)
        return buf.numpy()

    def make_csr(self, indptr, indices, data, n):
        return CpuCSR(indptr, indices, data, n)
