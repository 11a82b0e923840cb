"""Prior: kernel / mean dispatch (reference: fvgp/gp_prior.py).

The reference evaluates `kernel(x, x, hps)` with numpy on the host (default kernel:
gp_prior.py:376-400) or, under gp2Scale, block by block on dask workers
(gp_prior.py:324-370 -> gp2Scale_covariance.py).  Here every kernel evaluation ends in one
fused CUDA launch (dense) or a count+fill pair (gp2Scale) and the result stays on the
device for the factorisation; host copies are made only when a caller asks for
`prior.K` / `compute_prior_covariance_matrix`.
"""
import inspect
import warnings

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from . import kernels as K
from . import ops


class GPprior:
    def __init__(self, data, trainer, kernel=None, prior_mean_function=None, kernel_grad=None,
                 prior_mean_function_grad=None, gp2Scale_batch_size=10000, gp2Scale_distribution="blockwise"):
        self.data, self.trainer = data, trainer
        self.gp2Scale_batch_size = gp2Scale_batch_size          # kept for API parity: tiling is the kernel's own
        self.gp2Scale_distribution = gp2Scale_distribution
        if gp2Scale_distribution not in ("blockwise", "rowwise"):
            raise Exception(f"Unknown gp2Scale distribution `{gp2Scale_distribution}`.")
        self.default_kernel = kernel is None
        if kernel is None:
            if data.gp2Scale:
                warnings.warn("gp2Scale without a compactly supported kernel: using the anisotropic Wendland kernel.")
                kernel = K.wendland_anisotropic_gp2Scale_cpu          # gp_prior.py:49-50
            else:
                kernel = self._default_kernel
        self.kernel = kernel
        self.k_n_params = len(inspect.signature(kernel).parameters)   # gp_prior.py:61
        if self.k_n_params not in (3, 4):
            raise Exception("No valid kernel function signature")
        self.kernel_grad = kernel_grad
        self.mean_function = prior_mean_function
        self.mean_function_grad = prior_mean_function_grad
        self._K_host = None
        self.m = self.compute_mean(self.x_data, self.hyperparameters)

    # ---- shared state ----------------------------------------------------------------------
    @property
    def x_data(self):
        return self.data.x_data

    @property
    def y_data(self):
        return self.data.y_data

    @property
    def hyperparameters(self):
        return self.trainer.hyperparameters

    @property
    def args(self):
        return self.data.args

    # ---- kernel evaluation -----------------------------------------------------------------
    @staticmethod
    def _default_kernel(x1, x2, hyperparameters):
        """ARD Matern-3/2 (gp_prior.py:376-400), as a lazy fused expression."""
        hps = np.asarray(hyperparameters, dtype=np.float64)
        d = K.Distance(x1, x2, 1.0 / hps[1:1 + np.shape(x1)[1]])
        return K.Radial(d, L.K_MATERN32, 1.0, hps[0])

    def _call_kernel(self, x1, x2, hps):
        if self.k_n_params == 4:
            return self.kernel(x1, x2, hps, self.args)
        return self.kernel(x1, x2, hps)

    def device_KV(self, hps, V, V_dev=None):
        """K(x_data, x_data; hps) + diag(V) on the device, ready to factor.

        Returns ("dense", (buf, ld)) with the LOWER triangle filled, or ("sparse", DeviceCSR).  V_dev: the caller's
        resident device copy of a vector V (saves the upload)."""
        x = self.x_data
        n = len(x)
        Vd = V_dev if V_dev is not None else (L.to_dev(V) if (V is not None and np.ndim(V) == 1) else None)
        res = self._call_kernel(x, x, np.asarray(hps, dtype=np.float64))
        xd = self.data.x_device() if self.data.Euclidean else None
        if isinstance(res, K.Radial):
            same = res.dist.x1 is x and res.dist.x2 is x
            out = res.materialize(mode=L.FILL_LOWER, noise=Vd, x1_dev=xd if same else None, x2_dev=xd if same else None)
            kind, obj = "dense", out
        elif isinstance(res, K.SparseWendland):
            same = res.x1 is x and res.x2 is x
            kind, obj = "sparse", res.to_device_csr(noise=Vd, x1_dev=xd if same else None, x2_dev=xd if same else None)
        elif isinstance(res, K._Lazy):
            buf, ld = res.materialize(mode=L.FILL_FULL)
            if Vd is not None:
                buf[:, :n].diagonal().add_(Vd)
            kind, obj = "dense", (buf, ld)
        elif sp.issparse(res):
            KV = res.tocsr().astype(np.float64)
            if V is not None:
                KV = KV.copy()
                KV.setdiag(KV.diagonal() + V)
            KV.sort_indices()
            torch = L._torch()
            obj = ops.DeviceCSR(L.to_dev(KV.indptr, torch.int64), L.to_dev(KV.indices, torch.int32), L.to_dev(KV.data),
                                KV.shape)
            kind = "sparse"
        else:
            arr = np.asarray(res, dtype=np.float64)
            buf, ld = L.dev_matrix(n, n)
            buf[:, :n] = L.to_dev(arr)
            if Vd is not None:
                buf[:, :n].diagonal().add_(Vd)
            kind, obj = "dense", (buf, ld)
        if V is not None and np.ndim(V) == 2:                       # matrix-valued noise (gp_kv.py:662-664)
            if kind != "dense":
                raise Exception("matrix-valued noise needs a dense covariance")
            obj[0][:, :n].add_(L.to_dev(np.asarray(V, dtype=np.float64)))
        if self.data.gp2Scale and kind == "dense":
            raise Exception("gp2Scale needs a compactly supported kernel returning a sparse covariance "
                            "(use fvgp_b200.kernels.wendland_anisotropic_gp2Scale_cpu)")
        return kind, obj

    def compute_covariances(self, x1, x2, hps):
        """Host-visible k(x1, x2) (gp_prior.py:217-224): ndarray, or csr_matrix under gp2Scale."""
        res = self._call_kernel(x1, x2, np.asarray(hps, dtype=np.float64))
        if isinstance(res, K.SparseWendland):
            return res.tocsr()
        if isinstance(res, K._Lazy):
            if res.shape[0] == res.shape[1] and getattr(getattr(res, "dist", res), "same", False):
                buf, _ = res.materialize(mode=L.FILL_SYMMETRIC)
                return buf[:, :res.shape[1]].cpu().numpy()
            return res.to_host()
        return res

    def compute_prior_covariance_matrix(self, x, hps):
        """gp_prior.py:185-195."""
        return self.compute_covariances(x, x, hps)

    def compute_data_cross_covariance(self, x_pred, hps):
        """gp_prior.py:200-215."""
        return self.compute_covariances(self.x_data, x_pred, hps)

    @property
    def K(self):
        """Prior covariance at the current hyperparameters (host copy, made on first access)."""
        if self._K_host is None:
            self._K_host = self.compute_prior_covariance_matrix(self.x_data, self.hyperparameters)
        return self._K_host

    # ---- gradients ---------------------------------------------------------------------------
    def dk_dh(self, x1, x2, hps, direction=None):
        """dK/dtheta (gp_prior.py:236-240): (H,U,V), or (U,V) for one `direction` (ram_economy)."""
        hps = np.asarray(hps, dtype=np.float64)
        if self.kernel_grad is not None:
            if direction is None:
                return self.kernel_grad(x1, x2, hps)
            return self.kernel_grad(x1, x2, hps, direction)
        if self.default_kernel and not self.data.gp2Scale:
            g = ops.kgrad_dense_matern32(K._device_points(x1), K._device_points(x2), hps).cpu().numpy()
            return g if direction is None else g[direction]
        if direction is None:
            return np.stack([self._dkernel_dh(x1, x2, i, hps) for i in range(len(hps))])
        return self._dkernel_dh(x1, x2, direction, hps)

    def _dkernel_dh(self, x1, x2, direction, hps):
        """Central finite difference of a user kernel, eps = 1e-8 (gp_prior.py:438-447)."""
        eps = 1e-8
        hp, hm = np.array(hps, dtype=np.float64), np.array(hps, dtype=np.float64)
        hp[direction] += eps
        hm[direction] -= eps
        return (np.asarray(self.compute_covariances(x1, x2, hp)) - np.asarray(self.compute_covariances(x1, x2, hm))) \
            / (2.0 * eps)

    # ---- mean ----------------------------------------------------------------------------------
    def compute_mean(self, x, hps):
        """gp_prior.py:226-234, default = mean of all y entries (:449-458)."""
        if self.mean_function is not None:
            m = np.asarray(self.mean_function(x, hps), dtype=np.float64)
            assert np.ndim(m) == 1 and len(m) == len(x), "prior mean function returned the wrong shape"
            return m
        return np.full(len(x), np.mean(self.y_data))

    def dm_dh(self, x, hps):
        """gp_prior.py:242-246: zeros for the default mean, user gradient, or finite differences."""
        if self.mean_function is None:
            return np.zeros((len(hps), len(x)))
        if self.mean_function_grad is not None:
            return np.asarray(self.mean_function_grad(x, hps))
        gr = np.empty((len(hps), len(x)))
        for i in range(len(hps)):
            hp, hm = np.array(hps, dtype=np.float64), np.array(hps, dtype=np.float64)
            hp[i] += 1e-6
            hm[i] -= 1e-6
            gr[i] = (self.compute_mean(x, hp) - self.compute_mean(x, hm)) / 2e-6
        return gr

    # ---- state ---------------------------------------------------------------------------------
    def update_state_hyperparameters(self):
        self._K_host = None
        self.m = self.compute_mean(self.x_data, self.hyperparameters)

    def update_state_data(self):
        self._K_host = None
        self.m = self.compute_mean(self.x_data, self.hyperparameters)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_K_host"] = self.K if len(self.x_data) <= 20000 else None      # test_pickle requires prior.K to survive
        return state
