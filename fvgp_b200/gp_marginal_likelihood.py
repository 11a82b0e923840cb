"""Log marginal likelihood and its gradient (reference: fvgp/gp_marginal_likelihood.py).

log_likelihood            :137-179   L = -1/2 ( (y-m)^T KV^-1 (y-m) / r + log|KV| + n log 2 pi )
neg_log_likelihood_gradient :224-309 g_i = -1/2 ( b^T dK_i b - tr(KV^-1 dK_i) )

The reference forms dK/dtheta as H dense N x N arrays and obtains the traces through
H stacked LU solves (gp_lin_alg.py:1581-1626): (3H+3) N^2 doubles and H*(8/3) N^3 flops.
Here KV^-1 comes from the Cholesky factor on the tensor cores (2/3 N^3 flops, in place) and
tr(KV^-1 dK_i) - b^T dK_i b = sum((KV^-1 - b b^T) o dK_i) is reduced tile by tile with dK_i
regenerated in registers, so nothing of size H x N x N ever exists.
"""
import numpy as np

from . import _lib as L
from . import ops


class GPMarginalLikelihood:
    def __init__(self, data, prior, likelihood, trainer, kv):
        self.data, self.prior, self.likelihood, self.trainer, self.kv = data, prior, likelihood, trainer, kv
        self._warm_start_KVinvY = None

    @property
    def args(self):
        return self.data.args

    def _x0(self):
        if not bool(self.args.get("sparse_krylov_warm_start", False)):
            return None
        for cand in (self._warm_start_KVinvY, self.kv.KVinvY):
            if cand is not None and cand.shape == self.data.y_data.shape:
                return cand
        return None

    # ------------------------------------------------------------------------------------------
    def log_likelihood(self, hyperparameters=None):
        y = self.data.y_data
        if hyperparameters is None:
            m, KVinvY, logdet = self.prior.m, self.kv.KVinvY, self.kv.logdet_KV
        else:
            hps = np.asarray(hyperparameters, dtype=np.float64)
            V = self.likelihood.calculate_V(self.data.x_data, hps)
            m = self.prior.compute_mean(self.data.x_data, hps)
            try:
                ev = self.kv.evaluate(hps, V, m, x0=self._x0(), want_factor=False)
            except Exception as e:
                raise Exception(f"Linear algebra failed for hyperparameters {hyperparameters}: {e}") from e
            KVinvY, logdet = ev.KVinvY, ev.logdet
            if bool(self.args.get("sparse_krylov_warm_start", False)):
                self._warm_start_KVinvY = np.array(KVinvY, copy=True)
        y_mean = y - m[:, None]
        l1 = np.sum(y_mean * KVinvY) / y_mean.shape[1]
        return float(-0.5 * (l1 + logdet + len(y) * np.log(2.0 * np.pi)))

    def log_likelihood_variance(self):
        v = getattr(self.kv, "last_logdet_variance", None)
        return None if v is None else 0.25 * float(v)

    def neg_log_likelihood(self, hyperparameters=None):
        return -self.log_likelihood(hyperparameters=hyperparameters)

    # ------------------------------------------------------------------------------------------
    def neg_log_likelihood_gradient(self, hyperparameters=None, component=0):
        if self.data.gp2Scale:
            raise Exception("Can't compute neg_log_likelihood_gradient for gp2Scale")      # :240
        hps = np.asarray(self.trainer.hyperparameters if hyperparameters is None else hyperparameters,
                         dtype=np.float64)
        x = self.data.x_data
        n, H = len(x), len(hps)
        V = self.likelihood.calculate_V(x, hps)
        m = self.prior.compute_mean(x, hps)
        ev = self.kv.evaluate(hps, V, m, want_logdet=False)
        if ev.sharded is not None:
            return self._gradient_sharded(hps, ev, component)
        if ev.factor is None:
            return self._gradient_host(hps, V, ev, component)
        b_dev = ev.alpha_dev[component].contiguous()
        b = ev.KVinvY[:, component]
        ops.potri(ev.factor)                                   # lower(buf) <- lower(KV^-1)
        fused = self.prior.default_kernel and self.prior.kernel_grad is None
        if fused:
            traces = ops.kgrad_trace_matern32(self.data.x_device(), hps, ev.factor.buf, ev.factor.ld, b_dev)
        elif self.prior.kernel_grad is None:
            traces = self._fused_user_kernel_traces(hps, ev, b_dev)
            fused = traces is not None
        dV = self.likelihood.calculate_V_grad(x, hps)
        have_dV = np.any(dV != 0.0)
        if have_dV:
            assert np.ndim(dV) == 2, "noise gradient must be 2-d (one diagonal per hyperparameter) on the GPU path"
            wdiag = ev.factor.buf[:, :n].diagonal().cpu().numpy() - b * b
        dm = self.prior.dm_dh(x, hps)
        grad = np.zeros(H)
        for i in range(H):
            gm = float(-dm[i] @ b)
            if gm == 0.0:                                      # the reference's switch, :301-308
                if fused:
                    t = traces[i]
                else:
                    dK = self.prior.dk_dh(x, x, hps, direction=i) if self.data.ram_economy \
                        else self._dk_all(x, hps)[i]
                    t = ops.trace_sym_product(ev.factor.buf, ev.factor.ld, b_dev, L.to_dev(np.asarray(dK)))
                if have_dV:
                    t += float(np.sum(wdiag * dV[i]))
                grad[i] = 0.5 * t
            grad[i] += gm
        self._dk_cache = None
        return grad

    def _descriptor(self, hps):
        """(kind, [amp, inv_scale_1..D, length]) of a user kernel that folds into one fused radial expression of
        the fvgp.kernels names on (x_data, x_data); None otherwise."""
        from . import kernels as K
        x = self.data.x_data
        if not self.data.Euclidean:
            return None
        res = self.prior._call_kernel(x, x, np.asarray(hps, dtype=np.float64))
        if not (isinstance(res, K.Radial) and res.dist.x1 is x and res.dist.x2 is x and res.kind in ops.TRACE_KINDS):
            return None
        inv = np.broadcast_to(np.asarray(res.dist.inv_scale, dtype=np.float64), (x.shape[1],))
        return res.kind, np.concatenate([[res.amp], inv, [res.length]])

    def _descriptor_jacobian(self, hps):
        """(kind, p0, J): descriptor parameters p = (amp, inv_scale_1..D, length) at hps and J = dp/dtheta
        (len(p) x H).  Exact for the default kernel (amp = theta_0, inv_scale_i = 1 / theta_{1+i}); central differences
        of the DESCRIPTOR for user kernels composed of the fvgp.kernels names (evaluating the lazy expression costs
        nothing).  None when the kernel does not fold into one fused radial expression."""
        hps = np.asarray(hps, dtype=np.float64)
        d0 = self._descriptor(hps)
        if d0 is None:
            return None
        kind, p0 = d0
        H = len(hps)
        J = np.zeros((len(p0), H))
        dim = self.data.x_data.shape[1]
        if self.prior.default_kernel:
            J[0, 0] = 1.0
            for i in range(dim):
                J[1 + i, 1 + i] = -1.0 / hps[1 + i] ** 2
            return kind, p0, J
        for h in range(H):
            step = 1e-6 * max(abs(hps[h]), 1e-3)
            hp, hm = np.array(hps, dtype=np.float64), np.array(hps, dtype=np.float64)
            hp[h] += step
            hm[h] -= step
            dp, dm = self._descriptor(hp), self._descriptor(hm)
            if dp is None or dm is None or dp[0] != kind or dm[0] != kind:
                return None
            J[:, h] = (dp[1] - dm[1]) / (2.0 * step)
        return kind, p0, J

    def _fused_user_kernel_traces(self, hps, ev, b_dev):
        """tr-terms sum_ij (KV^-1 - b b^T)_ij dK_ij/dtheta_h for a user kernel composed of fvgp.kernels names, without
        materialising dK (the reference forms it by finite differences, gp_prior.py:438-447: H dense N x N arrays).
        The kernel evaluates the traces against the descriptor's parameters p = (amp, inv_scale, length); the chain
        rule with dp/dtheta (_descriptor_jacobian) gives the traces against theta."""
        dj = self._descriptor_jacobian(hps)
        if dj is None:
            return None
        kind, p0, J = dj
        dim = self.data.x_data.shape[1]
        T = ops.kgrad_trace_radial(kind, self.data.x_device(), p0[0], p0[1:1 + dim], p0[-1], ev.factor.buf, ev.factor.ld,
                                   b_dev)
        return T @ J

    # ------------------------------------------------------------------------------------------
    # Population evaluation: what the optimisers actually do is evaluate MANY proposals --
    # differential-evolution generations (gp_training.py:60-80), multi-start / hgdl walkers, the H+1 gradients of the
    # finite-difference Hessian (:312-336).  One proposal at N ~ 1e3 is a latency-bound chain of ~70 small launches;
    # the population entry point overlaps the chains on concurrent streams with one host synchronisation.
    def population_supported(self, want_grad=False):
        """True when a set of proposals can go through ops.lml_population: dense Cholesky mode on one GPU, a kernel
        that folds into one fused radial expression, and noise / prior mean that do not depend on theta."""
        kv = self.kv
        if self.data.gp2Scale or kv.mode != "Chol" or not self.data.Euclidean:
            return False
        if self.likelihood.noise_function is not None or self.prior.mean_function is not None:
            return False
        V = self.likelihood.V
        if V is None or np.ndim(V) != 1 or self.data.y_data.shape[1] > 4 or kv._use_sharded(kv.mode, V):
            return False
        d = self._descriptor(self.trainer.hyperparameters) if want_grad else self._fill_descriptor(
            self.trainer.hyperparameters)
        return d is not None and (not want_grad or self.prior.kernel_grad is None)

    def _fill_descriptor(self, hps):
        """Like _descriptor, for every kind the fused K-fill knows (no gradient-trace requirement)."""
        from . import kernels as K
        x = self.data.x_data
        res = self.prior._call_kernel(x, x, np.asarray(hps, dtype=np.float64))
        if not (isinstance(res, K.Radial) and res.dist.x1 is x and res.dist.x2 is x):
            return None
        inv = np.broadcast_to(np.asarray(res.dist.inv_scale, dtype=np.float64), (x.shape[1],))
        return res.kind, np.concatenate([[res.amp], inv, [res.length]])

    def evaluate_population(self, thetas, with_gradient=False, component=0):
        """LML (and grad(-LML)) for every row of `thetas` (B, H).  Returns (lml (B,), grad (B, H) or None).
        Falls back to one-at-a-time evaluation when the population path does not apply."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        B, H = thetas.shape
        plan = None
        if B > 1 and self.population_supported(want_grad=with_gradient) and self.prior.default_kernel:
            # default kernel: amp = theta_0, inv_scale_i = 1 / theta_{1+i}, length = 1 -- no per-proposal kernel call
            dim = self.data.x_data.shape[1]
            plan = []
            for t in thetas:
                J = None
                if with_gradient:
                    J = np.zeros((dim + 2, H))
                    J[0, 0] = 1.0
                    J[np.arange(1, dim + 1), np.arange(1, dim + 1)] = -1.0 / t[1:1 + dim] ** 2
                plan.append((L.K_MATERN32, np.concatenate([[t[0]], 1.0 / t[1:1 + dim], [1.0]]), J))
        elif B > 1 and self.population_supported(want_grad=with_gradient):
            plan = []
            for t in thetas:
                dj = self._descriptor_jacobian(t) if with_gradient else self._fill_descriptor(t)
                if dj is None or dj[0] != (plan[0][0] if plan else dj[0]):
                    plan = None
                    break
                plan.append(dj)
        if plan is None:
            lml = np.array([self.log_likelihood(t) for t in thetas])
            grad = np.array([self.neg_log_likelihood_gradient(t, component=component) for t in thetas]) \
                if with_gradient else None
            return lml, grad
        x = self.data.x_data
        n, dim = x.shape
        y = self.data.y_data
        m = self.prior.compute_mean(x, thetas[0])
        y_mean = y - m[:, None]
        P = np.array([pl[1] for pl in plan])
        from . import kernels as K
        try:
            alpha, logdet, traces, info = ops.lml_population(
                plan[0][0], self.data.x_device(), P[:, 0], P[:, 1:1 + dim], P[:, -1], L.to_dev(self.likelihood.V),
                L.to_dev(np.ascontiguousarray(y_mean.T)), want_grad=with_gradient, component=component,
                bounds=K.point_bounds(x))
        except L.NativeLibraryError:
            raise
        bad = np.nonzero(info)[0]
        if bad.size:
            b = int(bad[0])
            raise Exception(f"Linear algebra failed for hyperparameters {thetas[b]}: "
                            f"{L.NonPositiveDefiniteError(int(info[b]), n)}")
        lml = np.empty(B)
        for b in range(B):
            KVinvY = alpha[b].T.copy()
            l1 = np.sum(y_mean * KVinvY) / y_mean.shape[1]
            lml[b] = -0.5 * (l1 + logdet[b] + n * np.log(2.0 * np.pi))
        grad = None
        if with_gradient:
            grad = np.array([0.5 * (traces[b] @ plan[b][2]) for b in range(B)])
        return lml, grad

    def log_likelihood_population(self, thetas):
        return self.evaluate_population(thetas, with_gradient=False)[0]

    def neg_log_likelihood_gradient_population(self, thetas, component=0):
        return self.evaluate_population(thetas, with_gradient=True, component=component)[1]

    def _gradient_sharded(self, hps, ev, component):
        """Block-cyclic multi-GPU path (fvgp_b200/sharded.py): distributed TRTRI + LAUUM, block traces, one
        all-reduce of H doubles.  Default kernel: analytic traces against theta.  User kernels composed of the
        fvgp.kernels names: traces against the fused descriptor (amp, inv_scale, length) times its Jacobian, as on one
        GPU (_fused_user_kernel_traces).  Same host-side switch on the mean gradient as the single-GPU path."""
        x = self.data.x_data
        if self.prior.kernel_grad is not None:
            raise Exception("the sharded dense gradient regenerates dK/dtheta on the device; a user kernel_function_grad "
                            "is not supported on this path")
        if np.any(self.likelihood.calculate_V_grad(x, hps) != 0.0):
            raise Exception("noise-function gradients are not supported on the sharded dense path")
        b = ev.KVinvY[:, component]
        if self.prior.default_kernel:
            traces = ev.sharded.gradient_traces(hps, ev.alpha_dev[component])
        else:
            dj = self._descriptor_jacobian(hps)
            if dj is None:
                raise Exception("the sharded dense gradient needs the default kernel or a kernel composed of the "
                                "fvgp_b200.kernels radial functions (squared exponential, exponential, Matern)")
            kind, p0, J = dj
            dim = x.shape[1]
            T = ev.sharded.gradient_traces(None, ev.alpha_dev[component], radial=(kind, p0[0], p0[1:1 + dim], p0[-1]))
            traces = T @ J
        dm = self.prior.dm_dh(x, hps)
        grad = np.zeros(len(hps))
        for i in range(len(hps)):
            gm = float(-dm[i] @ b)
            if gm == 0.0:                                      # the reference's switch, :301-308
                grad[i] = 0.5 * traces[i]
            grad[i] += gm
        return grad

    _dk_cache = None

    def _dk_all(self, x, hps):
        if self._dk_cache is None:
            try:
                self._dk_cache = self.prior.dk_dh(x, x, hps)
            except Exception as e:
                raise Exception("The gradient evaluation dK/dh + dNoise/dh was not successful. That normally means "
                                "the combination of ram_economy and definition of the gradient function is wrong.") from e
        return self._dk_cache

    def _gradient_host(self, hps, V, ev, component):
        """Custom-callable linalg modes: the reference algorithm on host arrays (user-supplied solver)."""
        x = self.data.x_data
        KV = self.kv.addKV(self.prior.compute_prior_covariance_matrix(x, hps), V)
        b = ev.KVinvY[:, component]
        dK = np.asarray(self.prior.dk_dh(x, x, hps))
        dV = self.likelihood.calculate_V_grad(x, hps)
        dm = self.prior.dm_dh(x, hps)
        obj = self.kv.mode[0](KV)
        grad = np.zeros(len(hps))
        for i in range(len(hps)):
            gm = float(-dm[i] @ b)
            if gm == 0.0:
                dKi = dK[i] + (np.diag(dV[i]) if np.ndim(dV[i]) == 1 else dV[i])
                grad[i] = -0.5 * (b @ dKi @ b - np.trace(np.asarray(self.kv.mode[1](obj, dKi))))
            grad[i] += gm
        return grad

    # ------------------------------------------------------------------------------------------
    def neg_log_likelihood_hessian(self, hyperparameters=None):
        """Forward finite difference of the gradient, eps = 1e-6 (:312-336); the H + 1 gradients are one population."""
        hps = np.asarray(self.trainer.hyperparameters if hyperparameters is None else hyperparameters, dtype=float)
        H = len(hps)
        out = np.zeros((H, H))
        pts = np.tile(hps, (H + 1, 1))
        for i in range(H):
            pts[1 + i, i] += 1e-6
        if self.population_supported(want_grad=True):
            g = self.neg_log_likelihood_gradient_population(pts)
        else:
            g = np.array([self.neg_log_likelihood_gradient(hyperparameters=p) for p in pts])
        for i in range(H):
            out[i, i:] = ((g[1 + i] - g[0]) / 1e-6)[i:]
        return out + out.T - np.diag(np.diag(out))

    def test_log_likelihood_gradient(self, hyperparameters, epsilon=1e-6):
        """Finite-difference vs analytic gradient of +LML (:338-364)."""
        thps = np.array(hyperparameters, dtype=float)
        fd = np.empty(len(thps))
        base = self.log_likelihood(hyperparameters=thps)
        for i in range(len(thps)):
            t = np.array(thps)
            t[i] += epsilon
            fd[i] = (self.log_likelihood(hyperparameters=t) - base) / epsilon
        return fd, -self.neg_log_likelihood_gradient(hyperparameters=thps)
