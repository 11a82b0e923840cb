"""Dense LML (+ gradient) with KV sharded over the GPUs of one box: 2-D block-cyclic Cholesky,
triangular solves, log-determinant, inverse and gradient traces.

Reference being replaced: the single-process numpy/scipy path
    K-fill            gp_prior.py:376-400 + gp_kv.py:640-669
    Cholesky          gp_lin_alg.py:237-269   (calculate_Chol_factor)
    solve / logdet    gp_lin_alg.py:289-360
    gradient traces   gp_marginal_likelihood.py:256-309 (through gp_lin_alg.py:1581-1626)
which cannot hold KV beyond one node's RAM (C5: N = 200 000 -> 320 GB); there is no multi-GPU
dense path in the reference.

Layout.  One process per GPU (torchrun), ranks arranged as a P x Q grid, rank = p*Q + q.  KV is cut
into nb x nb blocks; block (I, J) lives on rank (I mod P, J mod Q).  Only blocks on / below the
diagonal are stored: every rank keeps, per local block COLUMN J, one dense panel holding its
row blocks I >= J (row stride nb) -- the lower "staircase", so N = 200 000 costs 160 GB / (P*Q).

Factorisation (right-looking, per block column k):
    owner(k,k):   potrf of the diagonal block (recursive DMMA Cholesky of the single-GPU path)
    broadcast     L_kk (+ its tile inverses)            -> kept REPLICATED (solves, logdet)
    column q_k:   panel <- panel * L_kk^-T              (TRSM by tile inverses = GEMMs)
    broadcast     panel pieces, one per process row     -> every rank holds the whole block column
    everyone:     A_IJ -= L_Ik L_Jk^T on its own blocks (one DMMA GEMM per local block column)
No collective touches the O(N^3) work; the volume received per rank is 8*N^2/2 bytes in total.
Solves are blocked forward / backward substitutions on the replicated diagonal blocks with one
nb-vector all-reduce per block; the inverse (for the gradient) is the same pattern twice
(TRTRI, LAUUM) and the traces are local reductions followed by one all-reduce of H doubles.

The numerical kernels are reached through a small `LocalOps` object.  `CudaLocalOps` (below) is the
product: every method is one call into the C ABI (include/fvgp_b200.h).  The choreography is
tested on CPU with gloo by injecting a torch-CPU LocalOps that lives in tests/ only.
"""
import math

import numpy as np

from . import _lib as L


# --------------------------------------------------------------------------------------------------
# layout arithmetic (pure Python; no device)
# --------------------------------------------------------------------------------------------------
def choose_grid(world):
    """P x Q with P <= Q, as square as possible (2 -> 1x2, 4 -> 2x2, 8 -> 2x4)."""
    p = int(math.isqrt(world))
    while world % p:
        p -= 1
    return p, world // p


class BlockCyclicLayout:
    def __init__(self, n, nb, P, Q, rank):
        assert nb % 128 == 0 and n > 0
        self.n, self.nb, self.P, self.Q, self.rank = int(n), int(nb), int(P), int(Q), int(rank)
        self.p, self.q = divmod(rank, Q)
        self.nblk = (n + nb - 1) // nb

    def bsize(self, I):
        return self.nb if I < self.nblk - 1 else self.n - (self.nblk - 1) * self.nb

    def owner(self, I, J):
        return (I % self.P) * self.Q + (J % self.Q)

    def row_blocks(self, p=None):
        return list(range(self.p if p is None else p, self.nblk, self.P))

    def col_blocks(self, q=None):
        return list(range(self.q if q is None else q, self.nblk, self.Q))

    def mloc(self, p=None):
        """Number of matrix rows process row p owns."""
        return sum(self.bsize(I) for I in self.row_blocks(p))

    def lrow(self, I):
        """Local row offset of global row block I on its process row."""
        return (I // self.P) * self.nb

    def rows_from(self, I, p=None):
        """Local row offset of the first row block >= I owned by process row p (== mloc(p) if none)."""
        p = self.p if p is None else p
        lb = max(0, -(-(I - p) // self.P))
        return min(lb * self.nb, self.mloc(p))

    def global_rows(self, p=None):
        out = [np.arange(I * self.nb, I * self.nb + self.bsize(I)) for I in self.row_blocks(p)]
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)


# --------------------------------------------------------------------------------------------------
# the single-GPU pieces: one C-ABI call each
# --------------------------------------------------------------------------------------------------
class CudaLocalOps:
    """Every method enqueues one entry point of libfvgp_b200.so on torch's current stream."""
    device = "cuda"

    def __init__(self):
        self.lib = L.load()
        self.torch = L._torch()
        self._gemv_work = None
        self._trsv_work = None
        self._tri_work = None
        self._partials = None

    def empty(self, *shape):
        return self.torch.empty(shape, dtype=self.torch.float64, device="cuda")

    def zeros(self, *shape):
        return self.torch.zeros(shape, dtype=self.torch.float64, device="cuda")

    def upload(self, a):
        return L.to_dev(a)

    def tileinv_len(self, n):
        return int(self.lib.fvgp_chol_workspace_len(n))

    @staticmethod
    def _ld(t):
        return t.stride(0) if t.shape[0] > 1 else max(t.shape[1] + (t.shape[1] % 2), t.stride(0))

    def fill(self, kind, x1, x2, amp, inv_scale, length, noise, out, centre):
        """out (m x n view) <- k(x1, x2) [+ diag(noise)]."""
        _, inv_p = L.dvec(inv_scale)
        cp = None
        if centre is not None:
            _keep, cp = L.dvec(centre)
        L.check(self.lib.fvgp_kfill_dense(kind, L.FILL_FULL, L.ptr(x1), x1.shape[0], L.ptr(x2), x2.shape[0], x1.shape[1],
                                          float(amp), inv_p, cp, float(length), L.ptr(noise), L.ptr(out), self._ld(out),
                                          L.stream_ptr()), "fvgp_kfill_dense")

    def potrf(self, A, n, tileinv):
        """In-place lower Cholesky of the n x n view A; returns the 1-based failing pivot or 0."""
        info = self.torch.zeros(1, dtype=self.torch.int32, device="cuda")
        return L.check(self.lib.fvgp_potrf_lower(L.ptr(A), n, self._ld(A), L.ptr(tileinv), L.ptr(info), L.stream_ptr()),
                       "fvgp_potrf_lower")

    def potrf_enqueue(self, A, n, tileinv, info_slot):
        """The same without a host synchronisation: the status is left in info_slot (1-element int32 device view)."""
        L.check(self.lib.fvgp_potrf_lower_enqueue(L.ptr(A), n, self._ld(A), L.ptr(tileinv), L.ptr(info_slot),
                                                  L.stream_ptr()), "fvgp_potrf_lower_enqueue")

    def info_buffer(self, count):
        return self.torch.zeros(count, dtype=self.torch.int32, device="cuda")

    def trsm_rlt(self, B, m, Lf, n, tileinv):
        L.check(self.lib.fvgp_trsm_right_lower_t(L.ptr(B), self._ld(B), m, L.ptr(Lf), self._ld(Lf), n, L.ptr(tileinv),
                                                 L.stream_ptr()), "fvgp_trsm_right_lower_t")

    # Trailing updates of at least this many rows (and block edges of at least int8_min_block) go through the INT8-slice
    # GEMM on tcgen05 (fvgp_ozaki_gemm: 2-3x the DMMA rate at FP64-grade accuracy, csrc/ozaki.cu) while fvgp_set_ozaki
    # is on; everything else -- and every product the INT8 path refuses -- runs on the DMMA pipe.
    int8_min_rows = 4096
    int8_min_block = 1024
    INT8_NBLOCK = 4096

    def _gemm_int8(self, a_mn, b_mn, A, B, C, m, n, k, alpha, beta):
        slices = self.lib.fvgp_ozaki_slices()
        if slices <= 0:
            return False
        need = int(self.lib.fvgp_ozaki_gemm_work_bytes(int(a_mn), int(b_mn), m, n, k, slices, self.INT8_NBLOCK))
        if need <= 0:
            return False
        key = L.stream_ptr().value or 0                           # one scratch buffer per stream: updates on the main and on
        pool = self.__dict__.setdefault("_int8_work", {})      # the look-ahead stream may be in flight together
        buf = pool.get(key)
        if buf is None or buf.numel() < need:
            pool[key] = None
            buf = pool[key] = self.torch.empty(need + (need >> 3), dtype=self.torch.uint8, device="cuda")
        rc = self.lib.fvgp_ozaki_gemm(int(a_mn), int(b_mn), L.ptr(A), self._ld(A), L.ptr(B), self._ld(B), L.ptr(C),
                                      self._ld(C), m, n, k, float(alpha), int(beta == 0.0), slices, self.INT8_NBLOCK,
                                      L.ptr(buf), buf.numel(), L.stream_ptr())
        if rc == -100:
            raise RuntimeError("fvgp_ozaki_gemm failed after part of the block had been updated")
        self.int8_calls = getattr(self, "int8_calls", 0) + (rc == 0)
        return rc == 0

    def gemm(self, a_mn, b_mn, A, B, C, m, n, k, alpha, beta, flags=0):
        if (flags == 0 and m >= self.int8_min_rows and min(n, k) >= self.int8_min_block and abs(alpha) == 1.0
                and beta in (0.0, 1.0) and self._gemm_int8(a_mn, b_mn, A, B, C, m, n, k, alpha, beta)):
            return
        L.check(self.lib.fvgp_dgemm(int(a_mn), int(b_mn), L.ptr(A), self._ld(A), L.ptr(B), self._ld(B), L.ptr(C),
                                    self._ld(C), m, n, k, float(alpha), float(beta), int(flags), L.stream_ptr()),
                "fvgp_dgemm")

    def _work(self, name, length):
        buf = getattr(self, name)
        if buf is None or buf.numel() < length:
            buf = self.empty(int(length))
            setattr(self, name, buf)
        return buf

    def trtri(self, A, n, tileinv):
        work = self._work("_tri_work", self.lib.fvgp_potri_workspace_len(n))
        L.check(self.lib.fvgp_trtri_lower(L.ptr(A), n, self._ld(A), L.ptr(tileinv), L.ptr(work), L.stream_ptr()),
                "fvgp_trtri_lower")

    def lauum(self, A, n):
        work = self._work("_tri_work", self.lib.fvgp_potri_workspace_len(n))
        L.check(self.lib.fvgp_lauum_lower(L.ptr(A), n, self._ld(A), L.ptr(work), L.stream_ptr()), "fvgp_lauum_lower")

    def trsv(self, Lf, n, tileinv, b, transpose):
        work = self._work("_trsv_work", 2 * n)
        L.check(self.lib.fvgp_trsv_lower(L.ptr(Lf), n, self._ld(Lf), L.ptr(tileinv), L.ptr(b), int(transpose),
                                         L.ptr(work), L.stream_ptr()), "fvgp_trsv_lower")

    def gemv(self, transpose, A, m, n, alpha, x, y):
        """y += alpha * A x (transpose = 0) or y += alpha * A^T x (transpose = 1); A is an m x n view."""
        work = self._work("_gemv_work", self.lib.fvgp_gemv_work_len(m, n))
        L.check(self.lib.fvgp_gemv(int(transpose), L.ptr(A), self._ld(A), m, n, float(alpha), L.ptr(x), L.ptr(y),
                                   L.ptr(work), L.stream_ptr()), "fvgp_gemv")

    def logdet(self, Lf, n):
        import ctypes
        scratch = self._work("_trsv_work", 2 * n)
        out = ctypes.c_double()
        L.check(self.lib.fvgp_chol_logdet(L.ptr(Lf), n, self._ld(Lf), L.ptr(scratch), ctypes.byref(out), L.stream_ptr()),
                "fvgp_chol_logdet")
        return out.value

    def trace_block(self, x1, x2, theta, W, m, n, b1, b2, diag_rows, accum):
        partials = self._work("_partials", self.lib.fvgp_kgrad_block_partials_len(x1.shape[1]))
        assert np.size(theta) >= x1.shape[1] + 1, "theta needs a signal variance and one length scale per dimension"
        _, th = L.dvec(theta)
        L.check(self.lib.fvgp_kgrad_trace_block_matern32(L.ptr(x1), m, L.ptr(x2), n, x1.shape[1], th, L.ptr(W),
                                                         self._ld(W), L.ptr(b1), L.ptr(b2), int(diag_rows),
                                                         L.ptr(partials), L.ptr(accum), L.stream_ptr()),
                "fvgp_kgrad_trace_block_matern32")


    def trace_block_radial(self, kind, x1, x2, inv_scale, length, W, m, n, b1, b2, diag_rows, accum):
        """accum (dim + 1, device) += raw sums of the block trace for a fused radial family."""
        partials = self._work("_partials", self.lib.fvgp_kgrad_block_partials_len(x1.shape[1]))
        assert np.size(inv_scale) == x1.shape[1], "one inverse length scale per input dimension"
        _, inv = L.dvec(inv_scale)
        L.check(self.lib.fvgp_kgrad_trace_block_radial(int(kind), L.ptr(x1), m, L.ptr(x2), n, x1.shape[1], inv,
                                                       float(length), L.ptr(W), self._ld(W), L.ptr(b1), L.ptr(b2),
                                                       int(diag_rows), L.ptr(partials), L.ptr(accum), L.stream_ptr()),
                "fvgp_kgrad_trace_block_radial")


GEMM_LOWER, GEMM_KB_FROM_M, GEMM_KB_FROM_N, GEMM_KE_FROM_M = 1, 2, 4, 8


class _Overlap:
    """Two-stream choreography helper: a high-priority side stream next to torch's current stream on CUDA,
    plain sequential execution for LocalOps without streams (the CPU ops of the gloo tests)."""

    def __init__(self, ops):
        self.on = getattr(ops, "device", "cpu") == "cuda"
        if self.on:
            self.torch = ops.torch
            self.main = self.torch.cuda.current_stream()
            self.side_stream = self.torch.cuda.Stream(priority=-1)

    def side(self):
        import contextlib
        return self.torch.cuda.stream(self.side_stream) if self.on else contextlib.nullcontext()

    def _fence(self, src, dst):
        ev = self.torch.cuda.Event()
        ev.record(src)
        dst.wait_event(ev)

    def fence_main_to_side(self):
        if self.on:
            self._fence(self.main, self.side_stream)

    def fence_side_to_main(self):
        if self.on:
            self._fence(self.side_stream, self.main)


# --------------------------------------------------------------------------------------------------
# communication: torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests)
# --------------------------------------------------------------------------------------------------
class Comm:
    def __init__(self, P, Q):
        import torch.distributed as dist
        self.dist = dist
        self.P, self.Q = P, Q
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        assert self.world == P * Q, f"grid {P}x{Q} needs {P * Q} ranks, have {self.world}"
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.bytes_received = 0
        self.p, self.q = divmod(self.rank, Q)
        # sub-communicators of the grid; every rank creates every group, in the same order
        self.row_groups, self.col_groups = {}, {}
        if self.world > 1:
            for pp in range(P):
                ranks = [pp * Q + qq for qq in range(Q)]
                self.row_groups[pp] = dist.new_group(ranks) if len(ranks) > 1 else None
            for qq in range(Q):
                ranks = [pp * Q + qq for pp in range(P)]
                self.col_groups[qq] = dist.new_group(ranks) if len(ranks) > 1 else None

    def broadcast(self, t, src):
        if self.world > 1:
            self.dist.broadcast(t, src)
            if src != self.rank:
                self.bytes_received += t.numel() * t.element_size()

    def _group_broadcast(self, t, src, group):
        if self.world == 1 or group is None:
            return
        self.dist.broadcast(t, src, group=group)
        if src != self.rank:
            self.bytes_received += t.numel() * t.element_size()

    def row_broadcast(self, t, src):
        """Broadcast inside my process row (ranks (p, .)); src is a global rank of that row."""
        self._group_broadcast(t, src, self.row_groups.get(self.p))

    def col_broadcast(self, t, src):
        """Broadcast inside my process column (ranks (., q))."""
        self._group_broadcast(t, src, self.col_groups.get(self.q))

    def all_reduce(self, t):
        if self.world > 1:
            self.dist.all_reduce(t)


# --------------------------------------------------------------------------------------------------
# the distributed matrix
# --------------------------------------------------------------------------------------------------
class ShardedSPD:
    """Lower staircase of a symmetric positive definite matrix in 2-D block-cyclic layout."""

    def __init__(self, n, nb, grid, ops, comm):
        P, Q = grid
        self.lay = BlockCyclicLayout(n, nb, P, Q, comm.rank)
        self.ops, self.comm = ops, comm
        lay = self.lay
        self.mloc = lay.mloc()
        self.my_cols = lay.col_blocks()
        # one panel per local block column J: rows of my row blocks I >= J
        self.col_r0 = {J: lay.rows_from(J) for J in self.my_cols}
        self.cols = {J: ops.empty(max(self.mloc - self.col_r0[J], 0), nb) for J in self.my_cols}
        maxm = max(lay.mloc(pp) for pp in range(P))
        self.panel = [ops.empty(maxm, nb) for _ in range(P)]          # the current block column, per process row
        self.diag = None                                              # replicated L_kk (nblk, nb, nb)
        self.diag_tinv = None
        self.info_dev = None
        self.info = 0
        self.state = "empty"

    def local_bytes(self):
        return sum(t.numel() for t in self.cols.values()) * 8

    # ---- views ----------------------------------------------------------------------------------
    def block_rows(self, J, I_from):
        """View of block column J restricted to my rows of blocks >= I_from: (tensor view, rows)."""
        r = self.lay.rows_from(I_from) - self.col_r0[J]
        t = self.cols[J]
        return t[r:], t.shape[0] - r

    def block(self, I, J):
        """View of block (I, J) (must be mine)."""
        r = self.lay.lrow(I) - self.col_r0[J]
        return self.cols[J][r:r + self.lay.bsize(I), :self.lay.bsize(J)]

    # ---- assembly -------------------------------------------------------------------------------
    def fill(self, kind, x_dev, x_host, amp, inv_scale, length, noise_dev, centre=None):
        """Each rank evaluates exactly the blocks it stores (K-fill shards with no collective)."""
        lay, ops = self.lay, self.ops
        rows = lay.global_rows()
        self.x_rows = x_dev[ops.upload(rows).long()] if len(rows) else x_dev[:0]
        self.x_dev = x_dev
        for J in self.my_cols:
            bJ = lay.bsize(J)
            x2 = x_dev[J * lay.nb:J * lay.nb + bJ]
            view, m = self.block_rows(J, J)
            if m <= 0:
                continue
            r0 = self.col_r0[J]
            if J % lay.P == lay.p:                                   # the diagonal block is mine: noise goes here
                nz = noise_dev[J * lay.nb:J * lay.nb + bJ] if noise_dev is not None else None
                ops.fill(kind, self.x_rows[r0:r0 + bJ], x2, amp, inv_scale, length, nz, view[:bJ, :bJ], centre)
                if m > bJ:
                    ops.fill(kind, self.x_rows[r0 + bJ:], x2, amp, inv_scale, length, None, view[bJ:, :bJ], centre)
            else:
                ops.fill(kind, self.x_rows[r0:], x2, amp, inv_scale, length, None, view[:, :bJ], centre)
        self.state = "filled"

    # ---- Cholesky -------------------------------------------------------------------------------
    def _exchange_block_column(self, k, panel=None):
        """After the owners of block column k have finished their pieces: every rank receives all of them.
        panel[pp][:m_pp] = rows of process row pp with blocks > k, in pp's local order."""
        lay, comm = self.lay, self.comm
        panel = self.panel if panel is None else panel
        qk = k % lay.Q
        sizes = []
        for pp in range(lay.P):
            m_pp = lay.mloc(pp) - lay.rows_from(k + 1, pp)
            sizes.append(m_pp)
            if m_pp <= 0:
                continue
            buf = panel[pp][:m_pp]
            if lay.p == pp and lay.q == qk:
                view, _ = self.block_rows(k, k + 1)
                buf.copy_(view)
            comm.broadcast(buf, pp * lay.Q + qk)
        return sizes

    def _panel_path(self, k, panel):
        """The latency-bound part of step k: factor the diagonal block, replicate it, solve the block column
        against it and replicate the block column into `panel`.  Returns the local failing pivot (0 = none)."""
        lay, ops, comm = self.lay, self.ops, self.comm
        nb = lay.nb
        bk = lay.bsize(k)
        pk, qk = k % lay.P, k % lay.Q
        owner = pk * lay.Q + qk
        D, T = self.diag[k], self.diag_tinv[k]
        info = 0
        if comm.rank == owner:
            blk = self.block(k, k)
            D[:bk, :bk].copy_(blk)
            if self.info_dev is not None:                     # no host synchronisation per panel: flags read once at the end
                ops.potrf_enqueue(D, bk, T, self.info_dev[k:k + 1])
            else:
                st = ops.potrf(D, bk, T)
                if st > 0:
                    info = k * nb + st
            blk.copy_(D[:bk, :bk])
        comm.broadcast(D, owner)
        comm.broadcast(T, owner)
        if lay.q == qk:
            view, m = self.block_rows(k, k + 1)
            if m > 0:
                ops.trsm_rlt(view, m, D, bk, T)
        if k < lay.nblk - 1:
            self._exchange_block_column(k, panel)
        return info

    def _trailing_update(self, k, panel, cols):
        """A_IJ -= L_Ik L_Jk^T on my blocks of the block columns `cols` (all > k): one DMMA GEMM per column."""
        lay, ops = self.lay, self.ops
        bk = lay.bsize(k)
        r0 = lay.rows_from(k + 1)
        for J in cols:
            bJ = lay.bsize(J)
            C, m = self.block_rows(J, J)
            if m <= 0:
                continue
            a_off = lay.rows_from(J) - r0
            pj = J % lay.P
            b_off = lay.lrow(J) - lay.rows_from(k + 1, pj)
            Aop = panel[lay.p][a_off:a_off + m]
            Bop = panel[pj][b_off:b_off + bJ]
            ops.gemm(0, 0, Aop, Bop, C, m, bJ, bk, -1.0, 1.0, 0)

    def factor(self):
        """In-place lower Cholesky, right-looking with ONE STEP OF LOOK-AHEAD.  Returns 0 or the 1-based global
        index of the first bad pivot.

        Step k's trailing update runs on the main stream; as soon as block column k+1 has received its update the
        panel path of step k+1 (diagonal POTRF, broadcasts, TRSM, block-column exchange over NCCL) starts on a
        high-priority side stream and overlaps the rest of the update.  The received block columns are double
        buffered.  On CPU (gloo tests) the same order runs sequentially."""
        assert self.state == "filled"
        lay, ops, comm = self.lay, self.ops, self.comm
        nb, nblk = lay.nb, lay.nblk
        tl = ops.tileinv_len(nb)
        self.diag = ops.zeros(nblk, nb, nb)
        self.diag_tinv = ops.zeros(nblk, tl)
        if getattr(self, "panel_b", None) is None:
            self.panel_b = [ops.empty(*t.shape) for t in self.panel]
        bufs = (self.panel, self.panel_b)
        ov = _Overlap(ops)
        self.info_dev = ops.info_buffer(nblk) if hasattr(ops, "info_buffer") else None
        info_local = 0
        ov.fence_main_to_side()
        with ov.side():
            info_local = self._panel_path(0, bufs[0]) or info_local
        ov.fence_side_to_main()
        for k in range(nblk - 1):
            cur, nxt = bufs[k % 2], bufs[(k + 1) % 2]
            mine = [J for J in self.my_cols if J > k]
            self._trailing_update(k, cur, [J for J in mine if J == k + 1])      # block column k+1 first ...
            ov.fence_main_to_side()
            self._trailing_update(k, cur, [J for J in mine if J > k + 1])       # ... the rest is queued before the
            with ov.side():                                                     # host blocks in the next POTRF
                st = self._panel_path(k + 1, nxt)
            if st and not info_local:
                info_local = st
            ov.fence_side_to_main()
        if self.info_dev is not None:                         # first failing panel of this rank: ONE read for all panels
            st = self.info_dev.cpu().numpy()
            bad = np.nonzero(st)[0]
            if bad.size:
                info_local = int(bad[0]) * nb + int(st[bad[0]])
        flag = ops.zeros(1)
        flag[0] = float(info_local) if info_local else float("inf")
        if comm.world > 1:
            comm.dist.all_reduce(flag, op=comm.dist.ReduceOp.MIN)
        v = float(flag.item())
        self.info = 0 if math.isinf(v) else int(v)
        self.state = "factored"
        self._logdet = None
        return self.info

    def logdet(self):
        lay = self.lay
        if self.state != "factored":                                 # the inverse reuses self.diag
            return self._logdet
        self._logdet = float(sum(self.ops.logdet(self.diag[k], lay.bsize(k)) for k in range(lay.nblk)))
        return self._logdet

    # ---- solves -----------------------------------------------------------------------------------
    def solve(self, b_dev):
        """(L L^T)^-1 b for a replicated right-hand side (N,); returns the replicated solution."""
        assert self.state == "factored"
        lay, ops, comm = self.lay, self.ops, self.comm
        nb, nblk = lay.nb, lay.nblk
        z = b_dev.clone()
        acc = ops.zeros(max(self.mloc, 1))
        t = ops.zeros(nb)
        for k in range(nblk):                                        # L z = b
            bk = lay.bsize(k)
            t.zero_()
            if lay.p == k % lay.P:
                t[:bk].copy_(acc[lay.lrow(k):lay.lrow(k) + bk])
            comm.all_reduce(t)
            zk = z[k * nb:k * nb + bk]
            zk.sub_(t[:bk])
            ops.trsv(self.diag[k], bk, self.diag_tinv[k], zk, 0)
            if lay.q == k % lay.Q:
                view, m = self.block_rows(k, k + 1)
                if m > 0:
                    ops.gemv(0, view, m, bk, 1.0, zk, acc[lay.rows_from(k + 1):])
        x = z
        rows = lay.global_rows()
        x_loc = ops.zeros(max(self.mloc, 1))
        for k in range(nblk - 1, -1, -1):                            # L^T x = z
            bk = lay.bsize(k)
            t.zero_()
            if lay.q == k % lay.Q:
                view, m = self.block_rows(k, k + 1)
                if m > 0:
                    ops.gemv(1, view, m, bk, 1.0, x_loc[lay.rows_from(k + 1):], t)
            comm.all_reduce(t)
            xk = x[k * nb:k * nb + bk]
            xk.sub_(t[:bk])
            ops.trsv(self.diag[k], bk, self.diag_tinv[k], xk, 1)
            if lay.p == k % lay.P:
                x_loc[lay.lrow(k):lay.lrow(k) + bk].copy_(xk)
        del rows
        return x

    # ---- inverse (for the gradient): TRTRI then LAUUM, both in place ---------------------------------
    def _row_panel_pieces(self, k):
        """Blocks (k, j), j < k, of my block columns (only if I am in process row k mod P)."""
        lay = self.lay
        return [J for J in self.my_cols if J < k] if lay.p == k % lay.P else []

    def _gather_row_panel(self, k, everyone, slot=0):
        """rowbuf[q'][i] = block (k, j_i) for the i-th block column j_i < k of process column q'.
        everyone=False: only my own process column's pieces are fetched (TRTRI); True: all (LAUUM).
        slot: which of the two receive-buffer sets to use (the look-ahead double-buffers them)."""
        lay, ops, comm = self.lay, self.ops, self.comm
        nb, bk, pk = lay.nb, lay.bsize(k), k % lay.P
        out = {}
        for qq in range(lay.Q):
            cols_q = [J for J in lay.col_blocks(qq) if J < k]
            if not cols_q:
                continue
            if not everyone and qq != lay.q:
                continue
            buf = self._rowbuf(qq, len(cols_q), slot)
            if lay.p == pk and lay.q == qq:
                for i, J in enumerate(cols_q):                      # J < k <= nblk-1, so block J is full width
                    buf[i, :bk].copy_(self.block(k, J))
            if everyone:
                comm.broadcast(buf, pk * lay.Q + qq)
            else:
                comm.col_broadcast(buf, pk * lay.Q + qq)
            out[qq] = (buf, cols_q)
        return out

    def _rowbuf(self, qq, count, slot=0):
        store = self.__dict__.setdefault("_rowbufs", {})
        buf = store.get((slot, qq))
        if buf is None or buf.shape[0] < count:
            cap = len(self.lay.col_blocks(qq))
            buf = self.ops.zeros(cap, self.lay.nb, self.lay.nb)
            store[(slot, qq)] = buf
        return buf[:count]

    def invert(self):
        """Lower staircase <- lower staircase of (L L^T)^-1 (calculate_inv_from_chol, gp_lin_alg.py:1558).

        TRTRI then LAUUM, both right-looking over block rows with ONE STEP OF LOOK-AHEAD: every collective of step
        k+1 (row-panel gather, column-panel broadcast) is issued on the high-priority side stream as soon as its
        inputs exist and lands in the second set of receive buffers while the main stream runs step k's GEMMs.  In
        TRTRI the inputs of step k+1 are block row k+1, which step k updates FIRST (split GEMMs); in LAUUM block row
        k+1 still holds its TRTRI value until step k+1 overwrites it, so its gather has no producer to wait for."""
        assert self.state == "factored"
        lay, ops, comm = self.lay, self.ops, self.comm
        nb, nblk = lay.nb, lay.nblk
        ov = _Overlap(ops)
        if getattr(self, "panel_b", None) is None:
            self.panel_b = [ops.empty(*t.shape) for t in self.panel]
        colbufs = (self.panel[lay.p], self.panel_b[lay.p])
        # -------- TRTRI, right-looking.  Two parts of every step do not depend on the steps before it and are hoisted
        # out of the loop, where they left most ranks idle: (A) M_kk = L_kk^-1 of every diagonal block (replicated in
        # self.diag, which holds the inverses from here on), (B) the column panels C <- -C M_kk (block column k below
        # the diagonal is still the original L there: steps j < k only touch columns J < j) -- one local GEMM per
        # owned block column, all ranks busy.  The loop keeps what is truly sequential: row-panel gather, panel
        # broadcast along the process row, accumulation into the columns to the left, M_kk times the block row.
        # (A): L_kk and its tile inverses are replicated on every rank (solves, logdet), so the nblk independent diagonal
        # inversions are dealt round-robin over ALL ranks -- with the owner doing each one and broadcasting it before the
        # next started, this was a serial chain of nblk TRTRIs with every other rank waiting (0.2 s at C3 on 8 GPUs) --
        # and the results are exchanged afterwards.
        for k in range(nblk):
            if comm.rank == k % comm.world:
                ops.trtri(self.diag[k], lay.bsize(k), self.diag_tinv[k])
                self.diag[k].tril_()
        for k in range(nblk):
            bk = lay.bsize(k)
            Dk = self.diag[k]
            comm.broadcast(Dk, k % comm.world)
            if comm.rank == (k % lay.P) * lay.Q + (k % lay.Q):
                self.block(k, k).copy_(Dk[:bk, :bk])
        self.state = "inverting"                                     # self.diag no longer holds the factor
        for J in self.my_cols:
            bJ = lay.bsize(J)
            view, m = self.block_rows(J, J + 1)
            if m > 0:
                tmp = self.panel[lay.p][:m]
                ops.gemm(0, 1, view, self.diag[J], tmp, m, bJ, bJ, -1.0, 0.0, GEMM_KB_FROM_N)
                view[:, :bJ].copy_(tmp[:, :bJ])

        def fetch_trtri(k, slot):
            """Collectives of TRTRI step k: my process column's piece of block row k, my process row's piece of
            column panel k.  Returns (row pieces, column-panel buffer or None)."""
            rowp = self._gather_row_panel(k, everyone=False, slot=slot)
            buf = None
            if k < nblk - 1:
                m_p = self.mloc - lay.rows_from(k + 1)
                if m_p > 0:
                    buf = colbufs[slot][:m_p]
                    if lay.q == k % lay.Q:
                        view, _ = self.block_rows(k, k + 1)
                        buf.copy_(view)
                    comm.row_broadcast(buf, lay.p * lay.Q + k % lay.Q)
            return rowp, buf

        def trtri_update(k, rowp, buf, first_only):
            """C_iJ += P_ik R_kJ for my rows i > k of my block columns J < k.  first_only: just the rows of block
            k+1 (when they are mine) -- the producer of step k+1's row panel; otherwise everything after them."""
            if buf is None or lay.q not in rowp:
                return
            head = lay.bsize(k + 1) if (k + 1) % lay.P == lay.p else 0
            rbuf, cols_q = rowp[lay.q]
            for i, J in enumerate(cols_q):
                C, m = self.block_rows(J, k + 1)
                if m <= 0:
                    continue
                lo, hi = (0, min(head, m)) if first_only else (min(head, m), m)
                if hi > lo:
                    ops.gemm(0, 1, buf[lo:hi], rbuf[i], C[lo:hi], hi - lo, lay.bsize(J), lay.bsize(k), 1.0, 1.0, 0)

        cur = None
        if nblk > 1:
            ov.fence_main_to_side()
            with ov.side():
                cur = fetch_trtri(1, 1)
        for k in range(1, nblk):
            pk = k % lay.P
            D = self.diag[k]
            ov.fence_side_to_main()                                   # step k's panels have arrived
            rowp, buf = cur
            nxt = None
            if k < nblk - 1:
                trtri_update(k, rowp, buf, first_only=True)           # block row k+1 is final after this
                ov.fence_main_to_side()
                with ov.side():
                    nxt = fetch_trtri(k + 1, (k + 1) % 2)
                trtri_update(k, rowp, buf, first_only=False)
            # block row k: R <- M_kk * R
            if lay.p == pk and lay.q in rowp:
                rbuf, cols_q = rowp[lay.q]
                for i, J in enumerate(cols_q):
                    ops.gemm(0, 1, D, rbuf[i], self.block(k, J), lay.bsize(k), lay.bsize(J), lay.bsize(k), 1.0, 0.0,
                             GEMM_KE_FROM_M)
            cur = nxt
        ov.fence_side_to_main()
        D = ops.zeros(nb, nb)
        # -------- LAUUM: lower(M^T M), block row by block row
        arow = ops.zeros(nb, (max(self.mloc, 2) + 15) // 16 * 16)
        cur = None
        if nblk > 1:
            ov.fence_main_to_side()
            with ov.side():
                cur = self._gather_row_panel(1, everyone=True, slot=1)
        for k in range(nblk):
            bk = lay.bsize(k)
            pk, qk = k % lay.P, k % lay.Q
            owner = pk * lay.Q + qk
            D.copy_(self.diag[k])                                    # M_kk is already replicated (TRTRI pass A)
            rowp = {}
            if k > 0:
                ov.fence_side_to_main()
                rowp = cur
            if k + 1 < nblk:
                ov.fence_main_to_side()                               # the other buffer set is free once step k-1 is queued
                with ov.side():
                    cur = self._gather_row_panel(k + 1, everyone=True, slot=(k + 1) % 2)
            if k > 0:
                # A operand: blocks (k, i) for MY row blocks i < k, laid out in my local row order
                my_rows_lt = [I for I in lay.row_blocks() if I < k]
                for I in my_rows_lt:
                    rbuf, cols_q = rowp[I % lay.Q]
                    arow[:bk, lay.lrow(I):lay.lrow(I) + nb].copy_(rbuf[cols_q.index(I), :bk])
                r_end = lay.rows_from(k)
                if lay.q in rowp and r_end > 0:
                    rbuf, cols_q = rowp[lay.q]
                    for i, J in enumerate(cols_q):
                        rJ = self.col_r0[J]
                        m = r_end - rJ
                        if m <= 0:
                            continue
                        C = self.cols[J][:m]
                        ops.gemm(1, 1, arow[:, rJ:], rbuf[i], C, m, lay.bsize(J), bk, 1.0, 1.0, 0)
                # block row k: R <- M_kk^T R
                if lay.p == pk and lay.q in rowp:
                    rbuf, cols_q = rowp[lay.q]
                    for i, J in enumerate(cols_q):
                        ops.gemm(1, 1, D, rbuf[i], self.block(k, J), bk, lay.bsize(J), bk, 1.0, 0.0, GEMM_KB_FROM_M)
            if comm.rank == owner:
                ops.lauum(D, bk)
                self.block(k, k).copy_(D[:bk, :bk])
        ov.fence_side_to_main()
        self.state = "inverted"

    # ---- gradient traces ----------------------------------------------------------------------------
    def grad_traces(self, theta, b_dev, radial=None):
        """sum_ij (KV^-1 - b b^T)_ij dK_ij/dtheta_h for the default kernel, all-reduced (H,) ndarray.
        radial = (kind, amp, inv_scale, length): instead the traces against the descriptor parameters
        (amp, inv_scale_1..D, length) of a fused radial kernel, (D + 2,) ndarray (the single-GPU counterpart is
        fvgp_kgrad_trace_radial)."""
        assert self.state == "inverted"
        lay, ops, comm = self.lay, self.ops, self.comm
        H = self.x_dev.shape[1] + 1
        accum = ops.zeros(H)
        rows = lay.global_rows()
        b_rows = b_dev[ops.upload(rows).long()] if len(rows) else b_dev[:0]
        for J in self.my_cols:
            bJ = lay.bsize(J)
            view, m = self.block_rows(J, J)
            if m <= 0:
                continue
            r0 = self.col_r0[J]
            diag_rows = bJ if J % lay.P == lay.p else 0
            if radial is None:
                ops.trace_block(self.x_rows[r0:], self.x_dev[J * lay.nb:J * lay.nb + bJ], theta, view, m, bJ,
                                b_rows[r0:], b_dev[J * lay.nb:J * lay.nb + bJ], diag_rows, accum)
            else:
                ops.trace_block_radial(radial[0], self.x_rows[r0:], self.x_dev[J * lay.nb:J * lay.nb + bJ], radial[2],
                                       radial[3], view, m, bJ, b_rows[r0:], b_dev[J * lay.nb:J * lay.nb + bJ],
                                       diag_rows, accum)
        comm.all_reduce(accum)
        raw = accum.cpu().numpy()
        if radial is None:
            return raw
        _kind, amp, inv_scale, length = radial
        inv_scale = np.broadcast_to(np.asarray(inv_scale, dtype=np.float64), (H - 1,))
        out = np.zeros(H + 1)
        out[0] = raw[0]
        for i in range(H - 1):
            out[1 + i] = -(amp / inv_scale[i]) * raw[1 + i] if inv_scale[i] != 0.0 else 0.0
        out[H] = (amp / length) * float(np.sum(raw[1:]))
        return out


# --------------------------------------------------------------------------------------------------
# LML (+ gradient) of the default kernel on the sharded matrix
# --------------------------------------------------------------------------------------------------
def default_block(n, world):
    """Block edge: large enough for GEMM efficiency, small enough for load balance (>= ~12 blocks per grid row)."""
    P, Q = choose_grid(world)
    nb = 2048
    while nb > 256 and n / nb < 12 * max(P, Q):
        nb //= 2
    return nb


class ShardedDenseEvaluator:
    """LML (+ gradient) of a radial-kernel GP with KV block-cyclic over all ranks.

    x, y, noise are replicated host arrays (24 bytes per point); everything O(N^2) is sharded."""

    def __init__(self, x, y, noise, nb=None, grid=None, ops=None, comm=None):
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_initialized() else 1
        self.grid = tuple(grid) if grid is not None else choose_grid(world)
        self.comm = comm if comm is not None else Comm(*self.grid)
        self.ops = ops if ops is not None else CudaLocalOps()
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64).reshape(len(x), -1)
        self.noise = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        self.n = len(self.x)
        self.nb = int(nb) if nb is not None else default_block(self.n, world)
        self.x_dev = self.ops.upload(self.x)
        self.noise_dev = None if self.noise is None else self.ops.upload(self.noise)
        self.bounds = (self.x.min(axis=0), self.x.max(axis=0))
        self.A = None
        self.last = {}

    def _matrix(self):
        if self.A is None:
            self.A = ShardedSPD(self.n, self.nb, self.grid, self.ops, self.comm)
        return self.A

    def evaluate(self, kind, amp, inv_scale, length, mean, want_gradient_theta=None, noise=None):
        """Returns dict(lml, alpha (N, r) ndarray, logdet[, traces]).  want_gradient_theta: the default
        kernel's theta when its gradient traces are wanted.  noise: replaces the noise diagonal (noise functions
        of the hyperparameters, gp_likelihood.py:89-94)."""
        from . import ops as single
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float64)
            if self.noise is None or not np.array_equal(noise, self.noise):
                self.noise = noise
                self.noise_dev = self.ops.upload(noise)
        A = self._matrix()
        centre = single.fill_centre(kind, inv_scale, length, self.bounds)
        self._marks = [("start", self._event())]
        A.fill(kind, self.x_dev, self.x, amp, inv_scale, length, self.noise_dev, centre)
        self._marks.append(("fill", self._event()))
        info = A.factor()
        self._marks.append(("factor", self._event()))
        if info > 0:
            raise L.NonPositiveDefiniteError(info, self.n)
        ym = self.y - np.asarray(mean, dtype=np.float64).reshape(-1, 1)
        cols = []
        for c in range(ym.shape[1]):
            cols.append(A.solve(self.ops.upload(np.ascontiguousarray(ym[:, c]))))
        self._marks.append(("solve", self._event()))
        alpha = np.stack([c.cpu().numpy() for c in cols], axis=1)
        logdet = A.logdet()
        r = ym.shape[1]
        lml = float(-0.5 * (np.sum(ym * alpha) / r + logdet + self.n * np.log(2.0 * np.pi)))
        out = {"lml": lml, "alpha": alpha, "logdet": logdet, "alpha_dev": cols}
        if want_gradient_theta is not None:
            out["traces"] = self.gradient_traces(want_gradient_theta, cols[0])
        self.last = out
        return out

    def gradient_traces(self, theta, b_dev, radial=None):
        """sum_ij (KV^-1 - b b^T)_ij dK_ij/dtheta_h for the default kernel at the hyperparameters of the last
        evaluate() -- or, with radial = (kind, amp, inv_scale, length), the traces against that descriptor's
        parameters; inverts the factored matrix in place on first use."""
        A = self._matrix()
        marks = getattr(self, "_marks", None)
        if marks is not None:
            marks.append(("(host)", self._event()))
        if A.state == "factored":
            A.invert()
        if marks is not None:
            marks.append(("invert", self._event()))
        out = A.grad_traces(None if theta is None else np.asarray(theta, dtype=np.float64), b_dev, radial=radial)
        if marks is not None:
            marks.append(("traces", self._event()))
        return out

    def _event(self):
        """CUDA event on the current stream (None for the CPU ops of the gloo tests)."""
        if getattr(self.ops, "device", "cpu") != "cuda":
            return None
        ev = self.ops.torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def phase_seconds(self):
        """Device time of the phases of the LAST evaluation (+ gradient) on this rank, from CUDA events on the main
        stream: {"fill", "factor", "solve", "invert", "traces"}; synchronises."""
        marks = [m for m in getattr(self, "_marks", []) if m[1] is not None]
        if len(marks) < 2:
            return None
        self.ops.torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            if name != "(host)":
                out[name] = out.get(name, 0.0) + a.elapsed_time(b) * 1e-3
        return out

    def solve(self, b):
        """KV^-1 b for host right-hand sides (N,) or (N, r) against the factor of the last evaluate()."""
        A = self._matrix()
        assert A.state == "factored", "the factor was consumed by the inverse; evaluate() again"
        b2 = np.asarray(b, dtype=np.float64).reshape(self.n, -1)
        out = np.stack([A.solve(self.ops.upload(np.ascontiguousarray(b2[:, c]))).cpu().numpy()
                        for c in range(b2.shape[1])], axis=1)
        return out.reshape(np.shape(b))
