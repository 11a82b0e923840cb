"""Posterior mean / covariance (reference: fvgp/gp_posterior.py:139-288) -- SURVEY 8(f) #1.

Re-uses the hot-path kernels: rectangular fused K-fill for k(x_pred, x_data), the resident
Cholesky factor for the multi-right-hand-side solve (tensor-core TRSM recursion) and the
DMMA GEMM for k KV^-1 k^T.  Under gp2Scale k stays sparse (device CSR) for the mean.
"""
import warnings

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from . import kernels as K
from . import ops


class GPposterior:
    def __init__(self, data, prior, trainer, kv, likelihood):
        self.data, self.prior, self.trainer, self.kv, self.likelihood = data, prior, trainer, kv, likelihood
        self.x_out = None

    @property
    def input_set_dim(self):
        return self.data.index_set_dim if self.x_out is None else self.data.index_set_dim - 1

    def _checks(self, x_pred, x_out):
        assert isinstance(x_pred, (np.ndarray, list)), "wrong format in x_pred"
        if isinstance(x_pred, np.ndarray):
            assert np.ndim(x_pred) == 2, "wrong dim in x_pred, has to be 2-d"
            assert x_pred.shape[1] == self.input_set_dim, "wrong number of columns in x_pred"
        assert x_out is None or isinstance(x_out, (np.ndarray, list)), "wrong format in x_out"
        if isinstance(x_out, np.ndarray):
            assert np.ndim(x_out) == 1, "wrong dim in x_out, has to be 1-d"

    @staticmethod
    def cartesian_product(x, y):
        """Task-major product space (gp_posterior.py:586-606), vectorised."""
        assert isinstance(y, np.ndarray) and np.ndim(y) == 1, "x_out must be a 1-d np.ndarray"
        return np.column_stack([np.tile(x, (len(y), 1)), np.repeat(y, len(x))])

    def _cross(self, x_pred, hps):
        """k(x_pred, x_data) on the device: ("dense", (n_pred, n) tensor) or ("sparse", DeviceCSR).

        The kernel is called as kernel(x_data, x_pred, hps) like the reference (gp_posterior.py:151).  The lazy result
        is evaluated on the point sets the KERNEL chose (a user kernel may slice or warp its inputs); the resident
        device copy of x_data is substituted only when the expression really is on (x_data, x_pred)."""
        x = self.data.x_data
        res = self.prior._call_kernel(x, x_pred, np.asarray(hps, dtype=np.float64))
        if isinstance(res, K.SparseWendland):
            if res.x1 is x and res.x2 is x_pred:
                return "sparse", ops.wendland_csr(K._device_points(x_pred), self.data.x_device(), res.hps)
            kt = res.tocsr().T.tocsr()                    # (n_pred, n), canonical
            kt.sort_indices()
            torch = L._torch()
            return "sparse", ops.DeviceCSR(L.to_dev(kt.indptr, torch.int64), L.to_dev(kt.indices, torch.int32),
                                           L.to_dev(kt.data), kt.shape)
        if isinstance(res, K.Radial):                   # radial kernels are symmetric in their arguments
            if res.dist.x1 is x and res.dist.x2 is x_pred:
                buf, _ = ops.kfill(res.kind, K._device_points(x_pred), self.data.x_device(), res.amp,
                                   res.dist.inv_scale, res.length, bounds=res.dist.bounds())
                return "dense", buf[:, :len(x)]
            buf, _ = res.materialize()                    # (n, n_pred) on the kernel's own point sets
            return "dense", buf[:, :res.shape[1]].t().contiguous()
        if isinstance(res, K._Lazy):
            return "dense", res.to_device().t().contiguous()
        if sp.issparse(res):
            res = res.toarray()
        return "dense", L.to_dev(np.asarray(res, dtype=np.float64)).t().contiguous()

    def posterior_mean(self, x_pred, hyperparameters=None, x_out=None):
        if x_out is None:
            x_out = self.x_out
        self._checks(x_pred, x_out)
        y = self.data.y_data
        if hyperparameters is not None:
            hps = np.asarray(hyperparameters, dtype=np.float64)
            V = self.likelihood.calculate_V(self.data.x_data, hps)
            m = self.prior.compute_mean(self.data.x_data, hps)
            KVinvY = self.kv.evaluate(hps, V, m, want_logdet=False).KVinvY
        else:
            hps, KVinvY = self.trainer.hyperparameters, self.kv.KVinvY
        x_orig = x_pred.copy()
        if isinstance(x_out, np.ndarray):
            x_pred = self.cartesian_product(x_pred, x_out)
        kind, k = self._cross(x_pred, hps)
        alpha_t = L.to_dev(np.ascontiguousarray(KVinvY.T))            # (r, n)
        r = alpha_t.shape[0]
        if kind == "dense":
            out = L.dev_empty((len(x_pred), r + (r % 2)))
            ops.dgemm_nt(k, alpha_t, out[:, :r])
            A = out[:, :r].cpu().numpy()
        else:
            A = np.stack([ops.spmv(k, alpha_t[c].contiguous()).cpu().numpy() for c in range(r)], axis=1)
        mean = self.prior.compute_mean(x_pred, hps)[:, None] + A
        if isinstance(x_out, np.ndarray):
            mean_re = mean.reshape(len(x_orig), len(x_out), order="F")
        else:
            mean_re = mean
        if y.shape[1] == 1 and not isinstance(x_out, np.ndarray):
            return {"x": x_orig, "m(x)": np.squeeze(mean_re), "m(x)_flat": np.squeeze(mean), "x_pred": x_pred}
        if y.shape[1] == 1:
            return {"x": x_orig, "m(x)": mean_re, "m(x)_flat": np.squeeze(mean), "x_pred": x_pred}
        return {"x": x_orig, "m(x)": mean_re, "m(x)_flat": mean, "x_pred": x_pred}

    def posterior_covariance(self, x_pred, x_out=None, variance_only=False, add_noise=False):
        if x_out is None:
            x_out = self.x_out
        self._checks(x_pred, x_out)
        x_orig = x_pred.copy()
        if isinstance(x_out, np.ndarray):
            x_pred = self.cartesian_product(x_pred, x_out)
        hps = self.trainer.hyperparameters
        npred, n = len(x_pred), len(self.data.x_data)
        kk = self.prior.compute_covariances(x_pred, x_pred, hps)
        kk = kk.toarray() if sp.issparse(kk) else np.asarray(kk)
        kind, k = self._cross(x_pred, hps)
        if kind == "dense" and self.kv.state.factor is not None:
            pad = max(5, npred) - npred                                # >4 rows selects the GEMM-based TRSM path
            rhs = L.dev_empty((npred + pad, n + (n % 2)))
            rhs.zero_()
            rhs[:npred, :n] = k
            if self.kv.state.factor.inverted:
                self.kv._refresh()
            lib = L.load()
            work = L.dev_empty((2 * n,))
            f = self.kv.state.factor
            L.check(lib.fvgp_potrs_lower(L.ptr(f.buf), n, f.ld, L.ptr(f.tileinv), L.ptr(rhs), npred + pad,
                                         rhs.stride(0), L.ptr(work), L.stream_ptr()), "fvgp_potrs_lower")
            prod = L.dev_empty((npred, npred + (npred % 2)))
            ops.dgemm_nt(k, rhs[:npred, :n], prod[:, :npred])
            S = kk - prod[:, :npred].cpu().numpy()
        else:
            kh = k.to_scipy().toarray() if kind == "sparse" else k.cpu().numpy()      # (npred, n)
            S = kk - kh @ self.kv.solve(kh.T)
        v = np.array(np.diag(S))
        if np.any(v < -0.0001):
            warnings.warn("Negative variances encountered. That normally means that the model is unstable. "
                          "Rethink the kernel definition, add more noise to the data, or double check the "
                          "hyperparameter optimization bounds.")
        if np.any(v < 0.0):
            v[v < 0.0] = 0.0
            np.fill_diagonal(S, v)
        if add_noise:
            noise = np.asarray(self.likelihood.calculate_V(x_pred, hps)) if self.likelihood.noise_function else None
            if noise is None:
                warnings.warn("Noise could not be added, you did not provide a noise callable at initialization")
            elif np.ndim(noise) == 1:
                v, S = v + noise, S + np.diag(noise)
            else:
                v, S = v + np.diag(noise), S + noise
        if variance_only:
            S = None
        if isinstance(x_out, np.ndarray):
            v_re = v.reshape(len(x_orig), len(x_out), order="F")
            S_re = None if S is None else S.reshape(len(x_orig), len(x_out), len(x_orig), len(x_out),
                                                    order="F").transpose(0, 2, 1, 3)
        else:
            v_re, S_re = v, S
            if self.data.y_data.shape[1] > 1:
                v = np.tile(v[:, None], (1, self.data.y_data.shape[1]))
                v_re = np.tile(v_re[:, None], (1, self.data.y_data.shape[1]))
        return {"x": x_orig, "x_pred": x_pred, "v(x)": v_re, "S": S_re, "S_flat": S, "v_flat": v}
