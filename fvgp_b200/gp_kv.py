"""K+V state and linear-algebra mode dispatch (reference: fvgp/gp_kv.py).

Modes keep the reference's names (gp_kv.py:138-141).  What runs where:
  Chol / CholInv / Inv          device: fused K-fill -> DMMA Cholesky -> solves / logdet (/ inverse)
  sparseCG / sparseCGpre        device: Wendland CSR -> (block-Jacobi) PCG, SLQ log-determinant
  sparseMINRES / sparseMINRESpre  same device PCG (KV is SPD, CG is the method of choice; same tolerance keys)
  sparseLU / sparseSolve        host scipy SuperLU on the device-assembled CSR (exact reference path, out of the
                                hot-path scope per SURVEY section 2 row 6; used to pin parity)
  (f_factor, f_solve, f_logdet) user callables, called with host arrays exactly like the reference
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import _lib as L
from . import ops

_DENSE = ("Chol", "CholInv", "Inv")
_CG = ("sparseCG", "sparseMINRES")
_CGPRE = ("sparseCGpre", "sparseMINRESpre")
_LU = ("sparseLU", "sparseSolve")


def resolve_gp2scale_linalg_mode(mode, args):
    """"sparseCGpre_<type>" -> ("sparseCGpre", args + preconditioner type) (gp_lin_alg.py:474-505)."""
    args = dict(args) if args is not None else {}
    for base in ("sparseCGpre", "sparseMINRESpre"):
        if mode.startswith(base + "_"):
            args["sparse_preconditioner_type"] = mode[len(base) + 1:]
            return base, args
    return mode, args


class Evaluation:
    """Result of one factor/solve/logdet pass; device objects are kept for the gradient / posterior."""
    __slots__ = ("KVinvY", "logdet", "factor", "csr", "alpha_dev", "mode", "info", "sharded", "serial")

    def __init__(self):
        self.KVinvY = self.logdet = self.factor = self.csr = self.alpha_dev = self.mode = None
        self.sharded = None          # ShardedDenseEvaluator holding the block-cyclic factor (multi-GPU dense path)
        self.serial = 0              # which evaluation of that evaluator this is (its matrix is reused in place)
        self.info = {}


class GPkv:
    allowed_modes = ["Chol", "CholInv", "Inv", "sparseMINRES", "sparseCG", "sparseLU", "sparseMINRESpre",
                     "sparseCGpre", "sparseMINRESpre_<type>", "sparseCGpre_<type>", "sparseSolve",
                     "a set of callables"]

    def __init__(self, data, prior, likelihood, linalg_mode=None):
        self.data, self.prior, self.likelihood = data, prior, likelihood
        self.last_logdet_variance = None
        self.last_logdet_info = {}
        if isinstance(linalg_mode, str):
            linalg_mode, self.data.args = resolve_gp2scale_linalg_mode(linalg_mode, self.data.args)
            if linalg_mode not in _DENSE + _CG + _CGPRE + _LU:
                raise Exception(f"No Mode. Choose from: {self.allowed_modes}")
        self.linalg_mode = linalg_mode
        self.mode = linalg_mode if linalg_mode is not None else ("Chol" if not data.gp2Scale else None)
        self.state = None
        self.KVinvY = None
        self.logdet_KV = None
        self._KVinv_host = None
        self._sparse_eval = None
        self.last_sharded_sparse_info = None
        self._refresh()

    # ---- properties mirroring the reference -----------------------------------------------------
    @property
    def args(self):
        return self.data.args

    @property
    def gp2Scale(self):
        return self.data.gp2Scale

    @property
    def Chol_factor(self):
        """Host copy of the lower factor (the reference stores it as an ndarray)."""
        f = self.state.factor if self.state is not None else None
        if f is None or f.inverted:
            return None
        return f.lower().cpu().numpy()

    @property
    def KVinv(self):
        if self.mode not in ("CholInv", "Inv"):
            return None
        if self._KVinv_host is None:
            f = self.state.factor
            if not f.inverted:
                ops.potri(f)
            low = np.tril(f.buf[:, :f.n].cpu().numpy())
            self._KVinv_host = low + np.tril(low, -1).T
        return self._KVinv_host

    @property
    def KV(self):
        """Host copy of K+V (csr_matrix under gp2Scale)."""
        if self.state is not None and self.state.csr is not None:
            return self.state.csr.to_scipy()
        return self.addKV(self.prior.K, self.likelihood.V)

    def _set_gp2Scale_mode(self, nnz):
        """gp_kv.py:182-188."""
        n = len(self.data.x_data)
        sparsity = float(nnz) / float(n ** 2)
        if self.linalg_mode is not None:
            return self.linalg_mode
        if n < 50001 and sparsity < 0.0001:
            return "sparseLU"
        if n < 2001 and sparsity >= 0.0001:
            return "Chol"
        return "sparseMINRES"

    # ---- the hot path -----------------------------------------------------------------------------
    def evaluate(self, hps, V, m, x0=None, want_logdet=True, want_factor=True):
        """K-fill (+V) -> factor -> KV^-1 (y-m) -> log|KV| for hyperparameters `hps`, stateless
        (compute_new_KVlogdet_KVinvY, gp_kv.py:574-631)."""
        # LML and gradient are separate calls in the reference API and each refactors KV
        # (gp_marginal_likelihood.py:158-165, :250-254); remember the last evaluation so that the
        # pair costs one K-fill + one Cholesky.
        key = (np.asarray(hps, dtype=np.float64).tobytes(), np.asarray(m).tobytes(),
               np.asarray(V).tobytes() if isinstance(V, np.ndarray) else id(V), len(self.data.x_data),
               self.data.generation, self._args_fingerprint())
        memo = getattr(self, "_memo", None)
        if memo is not None and memo[0] == key and x0 is None:
            old = memo[1]
            usable = old.factor is None or not old.factor.inverted
            if old.sharded is not None:
                usable = self.sharded_current(old)
            if usable and want_logdet and old.logdet is None and old.factor is not None:
                old.logdet = ops.chol_logdet(old.factor)
            if (usable or not want_factor) and (old.logdet is not None or not want_logdet):
                return old
        self._memo = None                                     # release the previous device buffers first
        ev = self._evaluate_uncached(hps, V, m, x0, want_logdet)
        self._memo = (key, ev)
        return ev

    def _dev_cached(self, name, arr):
        """Device copy of a host array that rarely changes between evaluations (the noise diagonal, y - m: 8 MB each
        at N = 1M, ~3.5 ms per pageable upload).  Keyed by content, so an in-place edit of the host array is seen."""
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        key = (arr.shape, arr.tobytes())
        store = self.__dict__.setdefault("_dev_cache", {})
        hit = store.get(name)
        if hit is None or hit[0] != key:
            hit = (key, L.to_dev(arr))
            store[name] = hit
        return hit[1]

    def _args_fingerprint(self):
        """Cheap identity of `args` for the memo key: a 4-argument kernel reads them, and the CG / SLQ keys change
        what an evaluation returns.  Scalars by value, everything else by object identity."""
        a = self.data.args
        if not a:
            return ()
        try:
            return tuple((k, v if isinstance(v, (bool, int, float, str, type(None))) else id(v))
                         for k, v in sorted(a.items(), key=lambda kv: str(kv[0])))
        except Exception:
            return (id(a),)

    def _evaluate_uncached(self, hps, V, m, x0, want_logdet):
        ev = Evaluation()
        y_mean = self.data.y_data - m[:, None]
        n, r = y_mean.shape
        mode = self.mode
        if callable_mode(mode):
            KV = self.addKV(self.prior.compute_prior_covariance_matrix(self.data.x_data, hps), V)
            obj = mode[0](KV)
            ev.KVinvY = np.asarray(mode[1](obj, y_mean)).reshape(y_mean.shape)
            ev.logdet = float(mode[2](obj)) if want_logdet else None
            ev.mode = mode
            return ev
        if self._use_sharded(mode, V):
            done = self._evaluate_sharded(ev, hps, V, m, y_mean)
            if done is not None:
                return done
        shard = self._sparse_shard(hps, V)
        if shard is not None:
            kind, obj = "sparse", shard
        else:
            V_dev = self._dev_cached("V", V) if (V is not None and np.ndim(V) == 1) else None
            kind, obj = self.prior.device_KV(hps, V, V_dev=V_dev)
        if kind == "sparse":
            mode = self._set_gp2Scale_mode(obj.nnz) if self.gp2Scale else (mode or "sparseCG")
            if mode in _DENSE:
                dense = L.to_dev(obj.to_scipy().toarray())
                buf, ld = L.dev_matrix(n, n)
                buf[:, :n] = dense
                kind, obj = "dense", (buf, ld)
        ev.mode = mode
        if kind == "dense":
            if mode not in _DENSE:
                raise Exception(f"Mode {mode} needs a sparse covariance (gp2Scale)")
            buf, ld = obj
            ev.factor = ops.potrf(buf, ld, n)
            rhs = L.to_dev(np.ascontiguousarray(y_mean.T))
            ops.potrs(ev.factor, rhs)
            ev.alpha_dev = rhs
            ev.KVinvY = rhs.cpu().numpy().T.copy()
            ev.logdet = ops.chol_logdet(ev.factor) if want_logdet else None
            return ev
        ev.csr = obj
        if mode in _LU:
            KV = obj.to_scipy().tocsc()
            lu = spla.splu(KV)
            ev.KVinvY = lu.solve(y_mean).reshape(y_mean.shape)
            ev.logdet = float(np.sum(np.log(np.abs(lu.L.diagonal()))) + np.sum(np.log(np.abs(lu.U.diagonal()))))
            ev.info["lu"] = lu
            return ev
        if mode not in _CG + _CGPRE:
            raise Exception(f"No mode: {mode}")
        rtol = float(self.args.get("sparse_cg_tol", self.args.get("cg_minres_tol", self.args.get("sparse_minres_tol", 1e-5))))
        maxiter = self.args.get("sparse_cg_maxiter", self.args.get("sparse_krylov_maxiter", None))
        precond = self._preconditioner(obj) if mode in _CGPRE else None
        sol = np.empty_like(y_mean)
        cols = []
        solver = self._sparse_eval.pcg if shard is not None else ops.pcg      # row-sharded PCG over NCCL | one GPU
        # The solve and the stochastic log-determinant only READ the assembled matrix and are bound by different
        # things (SpMV: HBM; the Lanczos SpMM: latency / L2 gathers), so on one GPU the log-determinant runs on a
        # side stream from a helper thread (ctypes releases the GIL inside the C calls) while this thread runs the CG.
        logdet_job = None
        if want_logdet and shard is None and self._overlap_logdet():
            logdet_job = self._start_logdet_job(obj, ev)
        try:
            for c in range(r):
                x0c = None if x0 is None else L.to_dev(np.ascontiguousarray(x0[:, c]))
                x, info, iters, relres = solver(obj, self._dev_cached(("y_mean", c), y_mean[:, c]), x0=x0c, rtol=rtol,
                                                maxiter=maxiter, precond=precond)
                ev.info.setdefault("cg_iters", []).append(iters)
                ev.info.setdefault("cg_relres", []).append(relres)
                cols.append(x)
                sol[:, c] = x.cpu().numpy()
        finally:
            if logdet_job is not None:
                ev.logdet = logdet_job()
        ev.alpha_dev = cols
        ev.KVinvY = sol
        if want_logdet and logdet_job is None:
            ev.logdet = self._random_logdet(obj, ev, sharded=shard is not None)
        return ev

    def _overlap_logdet(self):
        import os
        return os.environ.get("FVGP_SLQ_OVERLAP", "1") != "0" and not self.args.get("serial_logdet", False)

    def _start_logdet_job(self, csr, ev):
        """Run _random_logdet on a side stream in a helper thread; returns a function that joins it and hands back the
        estimate (re-raising whatever the thread raised)."""
        import threading
        torch = L._torch()
        main = torch.cuda.current_stream()
        side = getattr(self, "_side_stream", None)
        if side is None:
            side = self._side_stream = torch.cuda.Stream()
        side.wait_stream(main)                                  # the matrix is assembled on the main stream
        box = {}
        device = torch.cuda.current_device()

        def work():
            try:
                torch.cuda.set_device(device)
                with torch.cuda.stream(side):
                    box["value"] = self._random_logdet(csr, ev)
            except BaseException as e:                          # noqa: BLE001 -- handed to the caller's thread
                box["error"] = e
        th = threading.Thread(target=work, daemon=True)
        th.start()

        def join():
            th.join()
            main.wait_stream(side)
            if "error" in box:
                raise box["error"]
            return box["value"]
        return join

    # ---- multi-GPU gp2Scale: CSR row slabs + SLQ probes over the ranks of torch.distributed -----------
    def _sparse_shard(self, hps, V):
        """args["gp2Scale_sharded"] = True with more than one rank: assemble K + diag(V) with the rows sharded over
        all ranks (fvgp_b200/sharded_sparse.py) and return the replicated DeviceCSR; None otherwise.  Every rank must
        then call log_likelihood collectively with the same hyperparameters."""
        if not (self.gp2Scale and self.args.get("gp2Scale_sharded", False)) or callable_mode(self.mode):
            return None
        if self.mode is not None and self.mode not in _CG + _CGPRE:
            return None
        from . import kernels as K
        from . import sharded_sparse
        try:
            import torch.distributed as dist
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        except Exception:
            world = 1
        if world <= 1:
            return None
        x = self.data.x_data
        res = self.prior._call_kernel(x, x, np.asarray(hps, dtype=np.float64))
        if not (isinstance(res, K.SparseWendland) and res.x1 is x and res.x2 is x and V is not None and np.ndim(V) == 1):
            raise Exception("gp2Scale_sharded needs the anisotropic Wendland gp2Scale kernel on (x_data, x_data) and "
                            "vector noise")
        if getattr(self, "_sparse_eval", None) is None:
            self._sparse_eval = sharded_sparse.ShardedSparseEvaluator()
        csr, _rows = self._sparse_eval.assemble(self.data.x_device(), res.hps, self._dev_cached("V", V))
        self.last_sharded_sparse_info = dict(self._sparse_eval.info)
        return csr

    _BJACOBI_NAMES = (None, "", "default", "block_jacobi", "blockjacobi", "bjacobi", "block-jacobi", "jacobi")

    def _preconditioner(self, csr):
        """Block-Jacobi (32 x 32 dense inverse blocks, built and applied on the device) is the ONE preconditioner of
        this build; the reference's host-side ILU / IC / AMG / Schwarz family (gp_lin_alg.py:890-930) is out of scope
        (DESIGN.md section 7).  Any other `sparse_preconditioner_type` is answered with a warning, not silently."""
        want = self.args.get("sparse_preconditioner_type", None)
        key = want.lower() if isinstance(want, str) else want
        if key not in self._BJACOBI_NAMES and not getattr(self, "_warned_precond", False):
            import warnings
            warnings.warn(f"sparse_preconditioner_type={want!r} is not available on the B200 path; "
                          "using the device block-Jacobi preconditioner")
            self._warned_precond = True
        return ops.bjacobi(csr)

    # ---- multi-GPU dense path: KV block-cyclic over the ranks of torch.distributed ----------------
    def _use_sharded(self, mode, V):
        """args["dense_sharded"]: True -> always (also on one rank); "auto"/unset -> when torch.distributed has
        more than one rank AND KV + the inverse scratch would not fit one HBM.  Every rank must then call
        log_likelihood / neg_log_likelihood_gradient collectively with the same hyperparameters."""
        if self.gp2Scale or mode != "Chol" or V is None or np.ndim(V) != 1:
            return False
        want = self.args.get("dense_sharded", "auto")
        if want is False:
            return False
        if want is True:
            return True
        try:
            import torch.distributed as dist
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        except Exception:
            world = 1
        n = len(self.data.x_data)
        if world <= 1:
            return False
        hbm = float(L._torch().cuda.get_device_properties(L._torch().cuda.current_device()).total_memory)
        return 10.0 * n * n > 0.9 * hbm                          # 8 N^2 matrix + 2 N^2 inverse scratch

    def _evaluate_sharded(self, ev, hps, V, m, y_mean):
        from . import kernels as K
        from . import sharded
        x = self.data.x_data
        res = self.prior._call_kernel(x, x, np.asarray(hps, dtype=np.float64))
        if not (isinstance(res, K.Radial) and res.dist.x1 is x and res.dist.x2 is x):
            if self.args.get("dense_sharded") is True:
                raise Exception("dense_sharded needs the default kernel or a kernel composed of fvgp_b200.kernels "
                                "radial functions on get_(anisotropic_)distance_matrix")
            return None
        E = getattr(self, "_sharded_eval", None)
        if E is None or E.n != len(x) or x is not self._sharded_x:       # new data object -> new layout
            E = sharded.ShardedDenseEvaluator(x, self.data.y_data, V, nb=self.args.get("dense_sharded_block", None))
            self._sharded_eval, self._sharded_x, self._sharded_serial = E, x, 0
        E.y = np.ascontiguousarray(self.data.y_data, dtype=np.float64).reshape(len(x), -1)
        out = E.evaluate(res.kind, res.amp, res.dist.inv_scale, res.length, m, noise=V)
        self._sharded_serial += 1
        ev.sharded, ev.serial, ev.mode = E, self._sharded_serial, "Chol"
        ev.KVinvY, ev.logdet, ev.alpha_dev = out["alpha"], out["logdet"], out["alpha_dev"]
        return ev

    def sharded_current(self, ev):
        """True while the evaluator's in-place matrix still belongs to evaluation `ev` and is un-inverted."""
        return ev.sharded is not None and ev.serial == self._sharded_serial and ev.sharded._matrix().state == "factored"

    def _random_logdet(self, csr, ev=None, sharded=False):
        """SLQ estimate with the reference's argument keys (gp_lin_alg.py:1103-1181).  sharded: the probes are
        split over the ranks (same probe stream, same estimate as on one GPU)."""
        a = self.args
        degree = int(a.get("random_logdet_lanczos_degree", 20))
        lo = int(a.get("random_logdet_min_num_samples", 10))
        hi = int(a.get("random_logdet_max_num_samples", 5000))
        rtol = float(a.get("random_logdet_error_rtol", 0.01))
        seed = int(a.get("random_logdet_seed", 0))
        if sharded:
            def draw(count, probe0):
                return self._sparse_eval.slq_logdet(csr, degree, count, seed, probe0=probe0)[2]
        else:
            def draw(count, probe0):
                return ops.slq_logdet(csr, degree=degree, probes=count, seed=seed, probe0=probe0)[2]
        samples = np.asarray(draw(lo, 0), dtype=np.float64)
        # imate's documented stopping rule: z_c * std / sqrt(ns) <= error_rtol * |mean| at confidence 0.95 (z_c = 1.96),
        # checked from min_num_samples on; until it holds (or max_num_samples is reached) the SAME probe stream is
        # extended by the number of samples the current variance asks for
        z95 = 1.959963984540054
        while len(samples) > 1 and len(samples) < hi:
            mean, std = float(samples.mean()), float(samples.std(ddof=1))
            if not (np.isfinite(std) and std > 0.0) or z95 * std / np.sqrt(len(samples)) <= rtol * abs(mean):
                break
            need = int(np.ceil((z95 * std / (rtol * abs(mean))) ** 2))
            need = int(min(hi, max(need, len(samples) + 1)))
            samples = np.concatenate([samples, draw(need - len(samples), len(samples))])
        est = float(samples.mean())
        var = float(samples.var(ddof=1) / len(samples)) if len(samples) > 1 else float("nan")
        self.last_logdet_variance = var
        self.last_logdet_info = {"variance": var, "num_samples_used": len(samples), "lanczos_degree": degree}
        return est

    # ---- reference-named entry points ----------------------------------------------------------------
    def compute_new_KVlogdet_KVinvY(self, K, V, m, x0=None, hps=None):
        """Reference signature takes K; ours regenerates it on the device from `hps`."""
        ev = self.evaluate(self.prior.hyperparameters if hps is None else hps, V, m, x0=x0)
        return ev.KVinvY, ev.logdet

    def compute_new_KVinvY(self, KV, m, x0=None, hps=None, V=None):
        ev = self.evaluate(self.prior.hyperparameters if hps is None else hps,
                           self.likelihood.V if V is None else V, m, x0=x0, want_logdet=False)
        return ev.KVinvY

    @staticmethod
    def addKV(K, V):
        """gp_kv.py:640-669 on host arrays (API parity; the device path fuses this into the fill)."""
        if sp.issparse(K):
            if sp.issparse(V):
                return K + V
            KV = K.copy().tocsr()
            KV.setdiag(K.diagonal() + V)
            return KV
        if sp.issparse(V):
            V = V.toarray()
        if np.ndim(V) == 2:
            return K + V
        KV = np.array(K, copy=True)
        np.fill_diagonal(KV, np.diag(K) + V)
        return KV

    def _refresh(self):
        """Recompute factorisation, KVinvY and logdet for the current state (gp_kv.py:404-423)."""
        self._KVinv_host = None
        ev = self.evaluate(self.prior.hyperparameters, self.likelihood.V, self.prior.m)
        if self.gp2Scale or self.mode is None:
            self.mode = ev.mode
        self.state = ev
        self.KVinvY = ev.KVinvY
        self.logdet_KV = ev.logdet

    def update_state(self, appended_from=None):
        """Refresh after a hyperparameter or data change.  appended_from = the previous number of points when rows
        were APPENDED (update_gp_data(append=True), gp.py:689-749): the stored Cholesky factor is then extended by a
        bordered update instead of being recomputed (the reference does N sequential rank-1 updates,
        gp_lin_alg.py:1310-1477, gp_kv.py:462-508)."""
        if appended_from is not None and self._append_refresh(int(appended_from)):
            return
        self._memo = None                       # never hand a previous data set's factor to the new state
        self._refresh()

    def _append_refresh(self, n_old):
        """Bordered Cholesky: with a = n_old rounded down to the 128-row tile grid,
              rows [a, n) of K+V are (re)filled,  L21 = K21 L11^-T  (tensor-core TRSM),
              S = K22 - L21 L21^T (SYRK),  L22 = chol(S)  (POTRF with tile index offset a / 128),
        O(N^2 m) flops instead of O(N^3).  Returns False when the fast path does not apply."""
        from . import kernels as K
        ev0 = self.state
        x = self.data.x_data
        n = len(x)
        V = self.likelihood.V
        a = (n_old // 128) * 128
        if (self.gp2Scale or self.mode != "Chol" or ev0 is None or ev0.factor is None or ev0.factor.inverted
                or ev0.factor.n != n_old or a < 128 or n <= n_old or V is None or np.ndim(V) != 1 or len(V) != n
                or not self.data.Euclidean):
            return False
        hps = np.asarray(self.prior.hyperparameters, dtype=np.float64)
        res = self.prior._call_kernel(x, x, hps)
        if not (isinstance(res, K.Radial) and res.dist.x1 is x and res.dist.x2 is x):
            return False
        lib, torch = L.load(), L._torch()
        old = ev0.factor
        new, ld = L.dev_matrix(n, n)
        new[:a, :a].copy_(old.buf[:a, :a])
        tileinv = L.dev_empty((int(lib.fvgp_chol_workspace_len(n)),))
        tiles_a = a // 128
        tileinv[:tiles_a * 128 * 128].copy_(old.tileinv[:tiles_a * 128 * 128])
        xd = self.data.x_device()
        Vd = L.to_dev(V)
        bounds = res.dist.bounds()
        m2 = n - a
        # rows [a, n): off-diagonal block against the kept columns, then the diagonal block with the noise
        ops.kfill(res.kind, xd[a:], xd[:a], res.amp, res.dist.inv_scale, res.length, mode=L.FILL_FULL,
                  out=(new[a:, :a], ld), bounds=bounds)
        ops.kfill(res.kind, xd[a:], xd[a:], res.amp, res.dist.inv_scale, res.length, noise=Vd[a:], mode=L.FILL_LOWER,
                  out=(new[a:, a:], ld), bounds=bounds)
        st = L.stream_ptr()
        L.check(lib.fvgp_trsm_right_lower_t(L.ptr(new[a:, :a]), ld, m2, L.ptr(new), ld, a, L.ptr(tileinv), st),
                "fvgp_trsm_right_lower_t")
        L.check(lib.fvgp_dgemm(0, 0, L.ptr(new[a:, :a]), ld, L.ptr(new[a:, :a]), ld, L.ptr(new[a:, a:]), ld, m2, m2, a,
                               -1.0, 1.0, 1, st), "fvgp_dgemm")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        status = L.check(lib.fvgp_potrf_lower(L.ptr(new[a:, a:]), m2, ld, L.ptr(tileinv[tiles_a * 128 * 128:]),
                                              L.ptr(info), st), "fvgp_potrf_lower")
        if status > 0:
            raise L.NonPositiveDefiniteError(a + status, n)
        ev = Evaluation()
        ev.mode = "Chol"
        ev.factor = ops.CholFactor(new, ld, n, tileinv)
        y_mean = self.data.y_data - self.prior.m[:, None]
        rhs = L.to_dev(np.ascontiguousarray(y_mean.T))
        ops.potrs(ev.factor, rhs)
        ev.alpha_dev = rhs
        ev.KVinvY = rhs.cpu().numpy().T.copy()
        ev.logdet = ops.chol_logdet(ev.factor)
        ev.info["appended_rows"] = n - n_old
        self._memo = None
        self._KVinv_host = None
        self.state, self.KVinvY, self.logdet_KV = ev, ev.KVinvY, ev.logdet
        return True

    def solve(self, b, x0=None):
        """KV^-1 b with the stored factorisation (gp_kv.py:671-700); b is (N,) or (N, r) on the host."""
        b = np.asarray(b, dtype=np.float64)
        b2 = b.reshape(len(b), -1)
        ev = self.state
        if callable_mode(self.mode):
            obj = self.mode[0](self.addKV(self.prior.K, self.likelihood.V))
            return np.asarray(self.mode[1](obj, b2)).reshape(b.shape)
        if ev.sharded is not None:
            if not self.sharded_current(ev):                    # the in-place matrix moved on: rebuild the state
                self._memo = None
                self._refresh()
                ev = self.state
            return ev.sharded.solve(b2).reshape(b.shape)
        if ev.factor is not None:
            if ev.factor.inverted:
                self._refresh()
                ev = self.state
            rhs = L.to_dev(np.ascontiguousarray(b2.T))
            ops.potrs(ev.factor, rhs)
            return rhs.cpu().numpy().T.reshape(b.shape)
        if "lu" in ev.info:
            return ev.info["lu"].solve(b2).reshape(b.shape)
        rtol = float(self.args.get("sparse_cg_tol", self.args.get("cg_minres_tol", 1e-5)))
        precond = self._preconditioner(ev.csr) if self.mode in _CGPRE else None
        out = np.empty_like(b2)
        for c in range(b2.shape[1]):
            x, _, _, _ = ops.pcg(ev.csr, L.to_dev(np.ascontiguousarray(b2[:, c])), rtol=rtol, precond=precond)
            out[:, c] = x.cpu().numpy()
        return out.reshape(b.shape)

    def solve_device(self, rhs_t):
        """In-place dense solve for (nrhs, N) device right-hand sides (posterior covariance)."""
        if self.state.sharded is not None:
            sol = self.solve(rhs_t.cpu().numpy().T)
            rhs_t.copy_(L.to_dev(np.ascontiguousarray(sol.T)))
            return rhs_t
        if self.state.factor.inverted:
            self._refresh()
        return ops.potrs(self.state.factor, rhs_t)

    def logdet(self):
        return self.logdet_KV

    def __getstate__(self):
        state = dict(self.__dict__)
        state["state"] = None                    # device buffers are rebuilt on first use after unpickling
        state["_memo"] = None
        state["_sharded_eval"] = state["_sharded_x"] = None
        state["_sparse_eval"] = None
        state["_dev_cache"] = {}
        state["_side_stream"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if self.state is None:
            try:
                self._refresh()
            except L.NativeLibraryError:
                pass


def callable_mode(mode):
    return isinstance(mode, (tuple, list)) and len(mode) == 3 and all(callable(f) for f in mode)
