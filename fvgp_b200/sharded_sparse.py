"""gp2Scale across the GPUs of one box ("gp2Scale fill" and "PCG / SLQ" rows).

Reference being replaced: the dask cluster of gp2Scale --
    gp_prior.py:301-322                 scatter of x to the workers
    gp2Scale_covariance.py:313-431      distributed_covariance: blockwise / ROWWISE tasks (row_strip_csr :173-200),
                                        assemble_row_strips :290-296 (vstack of the strips)
    gp_lin_alg.py:1213-1291             calculate_sparse_conj_grad (single process)
    gp_lin_alg.py:1103-1181             calculate_random_logdet (imate: probes are independent samples)

One process per GPU (torchrun), every rank holds x (24 bytes per point) and calls collectively:

  assemble   rows are cut into one slab per rank, slab edges on multiples of 32 rows.  Pass 1 counts the entries of
             equal slabs; the counts are all-gathered (8 bytes per row) and the slab edges are moved to equal NNZ
             ("balance by a count pass"; a slab whose edges moved is counted again, the fill kernel needs the count
             pass's per-chunk offsets).  The global indptr is the scan of the gathered counts, so every rank writes
             its strip straight into its final position of the full indices / data arrays (no offsets to fix up) and
             one in-place all-gather per array replicates the matrix -- canonical CSR, bit-identical to the
             single-GPU assembly because every row is produced by the same kernel from the same points.
  solve      row-sharded PCG inside the C ABI (fvgp_pcg_sharded): slab SpMV + fused vector kernels, two scalar
             all-reduces and one all-gather of the search direction per iteration over NCCL.
  logdet     SLQ probes are independent: rank r runs probes [p_r, p_{r+1}) of the SAME counter-based Rademacher stream
             on its replica of the matrix; the samples are all-gathered, so the estimate equals the single-GPU one.

Device work goes through a small ops object; `CudaSparseOps` is the product (every method = C-ABI calls), the
choreography is tested on CPU with gloo by injecting a numpy/torch-CPU stand-in that lives in tests/ only.
"""
import ctypes
import os
import time

import numpy as np

from . import _lib as L


# --------------------------------------------------------------------------------------------------
# slab arithmetic (pure; no device)
# --------------------------------------------------------------------------------------------------
def equal_slabs(n, world, align=32):
    """world + 1 row offsets, interior ones on multiples of `align`, as equal as that allows."""
    per = -(-n // world)
    per = -(-per // align) * align
    return [min(n, r * per) for r in range(world)] + [n]


def balanced_slabs(cum, n, world, align=32):
    """Row offsets that give every rank ~1/world of the entries.  cum: inclusive prefix sums of the row counts
    (length n, numpy or torch); offsets are rounded to multiples of `align` and kept monotone."""
    total = int(cum[-1]) if n else 0
    offs = [0]
    for r in range(1, world):
        target = total * r // world
        lo, hi = 0, n                                   # first row whose inclusive prefix exceeds the target
        while lo < hi:
            mid = (lo + hi) // 2
            if int(cum[mid]) > target:
                hi = mid
            else:
                lo = mid + 1
        cut = min(n, max(offs[-1], (lo + align // 2) // align * align))
        offs.append(cut)
    return offs + [n]


def split_probes(count, world):
    """world + 1 offsets cutting `count` probes into contiguous, nearly equal shares."""
    return [count * r // world for r in range(world + 1)]


# --------------------------------------------------------------------------------------------------
# device pieces: C-ABI calls only
# --------------------------------------------------------------------------------------------------
def _loaded_nccl_path():
    """Path of the NCCL shared object already mapped into this process (the one torch ships), or None."""
    try:
        with open("/proc/self/maps") as fh:
            for line in fh:
                if "libnccl" in line and ".so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    return None


class CudaSparseOps:
    device = "cuda"

    def __init__(self):
        self.lib = L.load()
        self.torch = L._torch()
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self._comm = None

    # ---- communicator: our own NCCL communicator, unique id distributed through torch.distributed ------------
    def comm(self):
        if self._comm is None:
            lib, torch = self.lib, self.torch
            path = _loaded_nccl_path()
            L.check(lib.fvgp_nccl_attach(path.encode() if path else None), "fvgp_nccl_attach")
            ident = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = (ctypes.c_ubyte * 128)()
                L.check(lib.fvgp_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)), "fvgp_comm_unique_id")
                ident = torch.tensor(list(buf), dtype=torch.uint8)
            ident = ident.cuda()
            if self.world > 1:
                self.dist.broadcast(ident, 0)
            raw = bytes(ident.cpu().tolist())
            out = ctypes.c_void_p()
            L.check(lib.fvgp_comm_create(ctypes.c_char_p(raw), self.rank, self.world, ctypes.byref(out)), "fvgp_comm_create")
            self._comm = out
        return self._comm

    def close(self):
        if self._comm is not None:
            self.lib.fvgp_comm_destroy(self._comm)
            self._comm = None

    # ---- assembly ---------------------------------------------------------------------------------------------
    def aabb(self, x):
        from . import ops
        return ops.wendland_aabb(x)

    def zeros_i64(self, n):
        return self.torch.zeros(max(int(n), 1), dtype=self.torch.int64, device="cuda")

    def count_slab(self, x, boxes, row0, nrows, theta, counts):
        """counts[row0:row0+nrows] <- entries per row of the slab; returns the per-chunk scratch the fill needs."""
        lib, torch = self.lib, self.torch
        n, dim = x.shape
        chunk = torch.empty(int(lib.fvgp_wendland_chunk_len(max(nrows, 1), n)), dtype=torch.int32, device="cuda")
        if nrows > 0:
            _, th = L.dvec(theta)
            L.check(lib.fvgp_wendland_csr_count(L.ptr(x[row0:]), nrows, L.ptr(boxes[(row0 // 32) * 2 * dim:]), L.ptr(x), n,
                                                L.ptr(boxes), dim, th, L.ptr(counts[row0:]), L.ptr(chunk), None,
                                                L.stream_ptr()), "fvgp_wendland_csr_count")
        return chunk

    def scan(self, counts, n):
        lib, torch = self.lib, self.torch
        indptr = torch.empty(n + 1, dtype=torch.int64, device="cuda")
        scratch = torch.empty(int(lib.fvgp_scan_scratch_len(n)), dtype=torch.int64, device="cuda")
        total = ctypes.c_int64()
        L.check(lib.fvgp_exclusive_scan_i64(L.ptr(counts), n, L.ptr(indptr), L.ptr(scratch), ctypes.byref(total),
                                            L.stream_ptr()), "fvgp_exclusive_scan_i64")
        return indptr, int(total.value)

    def alloc_csr(self, nnz):
        torch = self.torch
        return L.dev_empty_rounded(nnz, torch.int32), L.dev_empty_rounded(nnz, torch.float64)

    def fill_slab(self, x, boxes, row0, nrows, theta, indptr, chunk, noise, indices, data):
        if nrows <= 0:
            return
        n, dim = x.shape
        _, th = L.dvec(theta)
        L.check(self.lib.fvgp_wendland_csr_fill(L.ptr(x[row0:]), nrows, L.ptr(boxes[(row0 // 32) * 2 * dim:]), L.ptr(x), n,
                                                L.ptr(boxes), dim, th, L.ptr(indptr[row0:]), L.ptr(chunk),
                                                L.ptr(noise[row0:]) if noise is not None else None, row0, L.ptr(indices),
                                                L.ptr(data), L.stream_ptr()), "fvgp_wendland_csr_fill")

    def allgatherv(self, t, elem_offsets):
        """In place: rank r owns elements [elem_offsets[r], elem_offsets[r+1]) of the 1-d tensor t."""
        if self.world == 1:
            return
        off = (ctypes.c_int64 * (self.world + 1))(*[int(o) * t.element_size() for o in elem_offsets])
        L.check(self.lib.fvgp_comm_allgatherv(self.comm(), L.ptr(t), off, L.stream_ptr()), "fvgp_comm_allgatherv")

    def prefix_at(self, indptr, rows):
        """Host values of indptr at the given rows (world + 1 numbers)."""
        idx = self.torch.as_tensor(list(rows), dtype=self.torch.int64, device="cuda")
        return [int(v) for v in indptr[idx].cpu().tolist()]

    def inclusive_cumsum_host_view(self, counts, n):
        """Object whose [i] gives the inclusive prefix sum of counts at row i (device cumsum, lazy host reads)."""
        cum = self.torch.cumsum(counts[:n], 0)

        class _View:
            def __getitem__(self_inner, i):
                return int(cum[i].item())
        return _View()

    # ---- solve / log-determinant -------------------------------------------------------------------------------
    def bjacobi(self, csr):
        from . import ops
        return ops.bjacobi(csr)

    def pcg(self, rows, csr, precond, b, x0, rtol, maxiter):
        lib, torch = self.lib, self.torch
        n = csr.shape[0]
        row0 = int(rows[self.rank])
        x = torch.zeros(n, dtype=torch.float64, device="cuda") if x0 is None else x0.clone().contiguous()
        work = L.dev_empty((int(lib.fvgp_pcg_sharded_work_len(n)),))
        iters, relres = ctypes.c_int(), ctypes.c_double()
        if maxiter is None:
            maxiter = 10 * n
        offs = (ctypes.c_int64 * (self.world + 1))(*[int(r) for r in rows])
        pre = None if precond is None else precond[(row0 // 32) * 1024:]
        st = L.check(lib.fvgp_pcg_sharded(self.comm(), n, offs, L.ptr(csr.indptr[row0:]), L.ptr(csr.indices), L.ptr(csr.data),
                                          L.ptr(pre), L.ptr(b), L.ptr(x), float(rtol), int(maxiter), L.ptr(work),
                                          ctypes.byref(iters), ctypes.byref(relres), L.stream_ptr()), "fvgp_pcg_sharded")
        return x, st, iters.value, relres.value

    def slq_samples(self, csr, degree, probe0, count, seed):
        from . import ops
        if count <= 0:
            return np.zeros(0)
        return ops.slq_logdet(csr, degree=degree, probes=count, seed=seed, probe0=probe0)[2]

    def gather_samples(self, mine, offsets):
        """All ranks receive the concatenation of every rank's samples (offsets: world + 1 probe offsets)."""
        if self.world == 1:
            return mine
        total = int(offsets[-1])
        buf = self.torch.zeros(total, dtype=self.torch.float64, device="cuda")
        a, b = int(offsets[self.rank]), int(offsets[self.rank + 1])
        if b > a:
            buf[a:b] = self.torch.as_tensor(mine, dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(buf)                      # disjoint supports: the sum is the concatenation
        return buf.cpu().numpy()

    def make_csr(self, indptr, indices, data, n):
        from . import ops
        return ops.DeviceCSR(indptr, indices, data, (n, n))


# --------------------------------------------------------------------------------------------------
# the evaluator
# --------------------------------------------------------------------------------------------------
class ShardedSparseEvaluator:
    """Assembly, solve and log-determinant of the gp2Scale system with rows / probes sharded over all ranks."""

    def __init__(self, ops=None):
        self.ops = ops if ops is not None else CudaSparseOps()
        self.rank, self.world = self.ops.rank, self.ops.world
        self.rows = None
        self.info = {}
        # FVGP_SHARDED_TIMING=1: wall-clock per phase with a device synchronisation at every mark (diagnosis only;
        # the marks serialise host and device)
        self.timing = {} if os.environ.get("FVGP_SHARDED_TIMING") == "1" else None
        self._t_last = None

    def _mark(self, name):
        if self.timing is None:
            return
        if getattr(self.ops, "device", "cpu") == "cuda":
            self.ops.torch.cuda.synchronize()
        now = time.perf_counter()
        if name is not None and self._t_last is not None:
            self.timing[name] = self.timing.get(name, 0.0) + (now - self._t_last)
        self._t_last = now

    def assemble(self, x_dev, theta, noise_dev):
        """K(x, x; theta) + diag(noise) as a replicated canonical CSR; returns (csr, row offsets)."""
        ops = self.ops
        n = int(x_dev.shape[0])
        self._mark(None)
        boxes = ops.aabb(x_dev)
        rows = equal_slabs(n, self.world)
        counts = ops.zeros_i64(n)
        r0, r1 = rows[self.rank], rows[self.rank + 1]
        chunk = ops.count_slab(x_dev, boxes, r0, r1 - r0, theta, counts)
        self._mark("count")
        ops.allgatherv(counts[:n], rows)
        self._mark("gather_counts")
        indptr, nnz = ops.scan(counts, n)
        self._mark("scan")
        recount = False
        if self.world > 1 and n >= 64 * self.world:
            at = ops.prefix_at(indptr, rows)
            share = [at[r + 1] - at[r] for r in range(self.world)]
            if max(share) > 1.03 * (nnz / self.world):          # equal rows are not equal work: move the edges
                new_rows = balanced_slabs(ops.inclusive_cumsum_host_view(counts, n), n, self.world)
                recount = new_rows != rows
                rows = new_rows
        if recount:
            r0, r1 = rows[self.rank], rows[self.rank + 1]
            scratch = ops.zeros_i64(n)
            chunk = ops.count_slab(x_dev, boxes, r0, r1 - r0, theta, scratch)
        self._mark("balance")
        indices, data = ops.alloc_csr(nnz)
        ops.fill_slab(x_dev, boxes, r0, r1 - r0, theta, indptr, chunk, noise_dev, indices, data)
        self._mark("fill")
        at = ops.prefix_at(indptr, rows)
        ops.allgatherv(indices, at)
        ops.allgatherv(data, at)
        self._mark("gather_csr")
        self.rows = rows
        self.info = {"rows": list(rows), "nnz": nnz, "nnz_per_rank": [at[r + 1] - at[r] for r in range(self.world)],
                     "rebalanced": bool(recount)}
        return ops.make_csr(indptr, indices, data, n), rows

    def pcg(self, csr, b, x0=None, rtol=1e-5, maxiter=None, precond=None):
        """Same contract as ops.pcg (scipy cg semantics); every rank returns the whole solution."""
        self._mark(None)
        out = self.ops.pcg(self.rows, csr, precond, b, x0, rtol, maxiter)
        self._mark("pcg")
        return out

    def slq_logdet(self, csr, degree, probes, seed, probe0=0):
        """(estimate, variance of the mean, samples): probes probe0 .. probe0 + probes - 1 of the single-GPU stream,
        split over the ranks."""
        offs = split_probes(int(probes), self.world)
        self._mark(None)
        mine = self.ops.slq_samples(csr, degree, int(probe0) + offs[self.rank], offs[self.rank + 1] - offs[self.rank], seed)
        self._mark("slq")
        samples = np.asarray(self.ops.gather_samples(mine, offs), dtype=np.float64)
        self._mark("gather_samples")
        est = float(samples.mean())
        var = float(samples.var(ddof=1) / len(samples)) if len(samples) > 1 else float("nan")
        return est, var, samples
