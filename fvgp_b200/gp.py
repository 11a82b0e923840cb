"""`GP`: drop-in for `fvgp.GP` on the B200 (reference: fvgp/gp.py:26-2267).

Constructor and method signatures are the reference's (gp.py:419-439, :781-801, :1310-1490).
Inputs and outputs are host numpy arrays; everything between them on the training hot
path -- covariance assembly, factorisation, solves, log-determinant, likelihood gradient --
runs in the CUDA library behind `fvgp_b200._lib` and stays on the device.
`compute_device` (default "cpu", as in the reference signature) and args["GPU_engine"] (any value,
"b200" included) are accepted for compatibility and do not select anything: the B200 is always
the compute device and there is no CPU fallback.  `dask_client` is accepted and ignored: the gp2Scale block loop
that needed it is a single pair of kernel launches here.
"""
import warnings

import numpy as np

from .gp_data import GPdata
from .gp_kv import GPkv
from .gp_likelihood import GPlikelihood
from .gp_marginal_likelihood import GPMarginalLikelihood
from .gp_posterior import GPposterior
from .gp_prior import GPprior
from .gp_training import GPtraining


def out_of_bounds(x, bounds):
    return bool(np.any(x < bounds[:, 0]) or np.any(x > bounds[:, 1]))


class GP:
    def __init__(self, x_data, y_data, init_hyperparameters=None, noise_variances=None, compute_device="cpu",
                 kernel_function=None, kernel_function_grad=None, noise_function=None, noise_function_grad=None,
                 prior_mean_function=None, prior_mean_function_grad=None, gp2Scale=False, dask_client=None,
                 gp2Scale_batch_size=10000, gp2Scale_distribution="blockwise", linalg_mode=None, ram_economy=False,
                 args=None):
        assert isinstance(noise_variances, np.ndarray) or noise_variances is None, "wrong format in noise_variances"
        assert init_hyperparameters is None or isinstance(init_hyperparameters, np.ndarray), \
            "wrong init_hyperparameters"
        assert isinstance(compute_device, str), "wrong format in compute_device"
        for fn, nm in ((kernel_function, "kernel_function"), (kernel_function_grad, "kernel_function_grad"),
                       (noise_function, "noise_function"), (noise_function_grad, "noise_function_grad"),
                       (prior_mean_function, "prior_mean_function"),
                       (prior_mean_function_grad, "prior_mean_function_grad")):
            assert callable(fn) or fn is None, f"wrong format in {nm}"
        assert len(x_data) == len(y_data), "x_data and y_data do not have the same lengths."
        self.data = GPdata(x_data, y_data, args=args, noise_variances=noise_variances, ram_economy=ram_economy,
                           gp2Scale=gp2Scale, compute_device=compute_device, dask_client=dask_client)
        hyperparameters = init_hyperparameters
        if self.data.Euclidean:
            if callable(kernel_function) or callable(prior_mean_function) or callable(noise_function):
                if init_hyperparameters is None:
                    raise Exception("You have provided callables for kernel, mean, or noise functions but no "
                                    "initial hyperparameters.")
            elif init_hyperparameters is None:
                hyperparameters = np.ones((self.index_set_dim + 1))
                warnings.warn("Hyperparameters initialized to a vector of ones.")
        if hyperparameters is None:
            raise Exception("'init_hyperparameters' not provided and could not be calculated. Please provide them ")
        self.trainer = GPtraining(self.data, hyperparameters)
        self.prior = GPprior(self.data, self.trainer, kernel=kernel_function,
                             prior_mean_function=prior_mean_function, kernel_grad=kernel_function_grad,
                             prior_mean_function_grad=prior_mean_function_grad,
                             gp2Scale_batch_size=gp2Scale_batch_size, gp2Scale_distribution=gp2Scale_distribution)
        self.likelihood = GPlikelihood(self.data, self.trainer, noise_function=noise_function,
                                       noise_function_grad=noise_function_grad)
        self.kv = GPkv(self.data, self.prior, self.likelihood, linalg_mode=linalg_mode)
        self.marginal_likelihood = GPMarginalLikelihood(self.data, self.prior, self.likelihood, self.trainer, self.kv)
        self.posterior = GPposterior(self.data, self.prior, self.trainer, self.kv, self.likelihood)

    # ---- properties (gp.py:576-647) ------------------------------------------------------------
    @property
    def x_data(self):
        return self.data.x_data

    @property
    def y_data(self):
        return self.data.y_data

    @property
    def noise_variances(self):
        return self.data.noise_variances

    @property
    def index_set_dim(self):
        return self.data.index_set_dim

    @property
    def input_set_dim(self):
        return self.data.index_set_dim

    @property
    def hyperparameters(self):
        return self.trainer.hyperparameters

    @property
    def K(self):
        return self.prior.K

    @property
    def m(self):
        return self.prior.m

    @property
    def V(self):
        return self.likelihood.V

    @property
    def args(self):
        return self.data.args

    @args.setter
    def args(self, value):
        self.data.args = value
        self.kv._memo = None                    # a 4-argument kernel / the CG and SLQ keys read them

    # ---- state changes ---------------------------------------------------------------------------
    def set_hyperparameters(self, hps):
        """gp.py:672-687."""
        self.trainer.hyperparameters = np.array(hps, dtype=np.float64)
        self.prior.update_state_hyperparameters()
        self.likelihood.update_state()
        self.kv.update_state()

    def get_hyperparameters(self):
        return self.trainer.hyperparameters

    def update_gp_data(self, x_new, y_new, noise_variances_new=None, append=True, gp_rank_n_update=None):
        """gp.py:689-749.  append=True extends the resident Cholesky factor by a bordered update on the device
        (GPkv._append_refresh, O(N^2 m)); otherwise the covariance is regenerated and refactored."""
        n_old = len(self.data.x_data)
        self.data.update(x_new, y_new, noise_variances_new, append=append)
        self.prior.update_state_data()
        self.likelihood.update_state()
        self.kv.update_state(appended_from=n_old if append else None)

    def _get_default_hyperparameter_bounds(self):
        """gp.py:754-775."""
        if not self.data.Euclidean:
            raise Exception("Please provide custom hyperparameter bounds to the training in the non-Euclidean setting")
        if len(self.hyperparameters) != self.index_set_dim + 1:
            raise Exception("Please provide custom hyperparameter_bounds when kernel, mean or noise"
                            " functions are customized")
        b = np.zeros((self.index_set_dim + 1, 2))
        b[0] = np.array([np.var(self.y_data) / 100., np.var(self.y_data) * 10.])
        for i in range(self.index_set_dim):
            rng_i = np.max(self.x_data[:, i]) - np.min(self.x_data[:, i])
            b[i + 1] = np.array([rng_i / 100., rng_i * 10.])
        return b

    def train(self, hyperparameter_bounds=None, objective_function=None, objective_function_gradient=None,
              objective_function_hessian=None, init_hyperparameters=None, method="mcmc", pop_size=20, tolerance=0.0001,
              max_iter=10000, mcmc_prior=None, mcmc_prop_distrs="normal", mcmc_args={}, bo_args=None,
              local_optimizer="L-BFGS-B", global_optimizer="genetic", constraints=(), dask_client=None, info=False,
              asynchronous=False, accept_only_if_improved=True):
        """gp.py:781-1141 (synchronous path)."""
        if asynchronous:
            raise Exception("asynchronous training needs dask actors; not available in fvgp_b200")
        if hyperparameter_bounds is None:
            hyperparameter_bounds = self._get_default_hyperparameter_bounds()
            warnings.warn("Default hyperparameter_bounds initialized because none were provided. "
                          "This will fail for custom kernel, mean, or noise functions")
        hyperparameter_bounds = np.asarray(hyperparameter_bounds, dtype=np.float64)
        lo, hi = hyperparameter_bounds[:, 0], hyperparameter_bounds[:, 1]
        if init_hyperparameters is None:
            init_hyperparameters = self.hyperparameters if not out_of_bounds(self.hyperparameters, hyperparameter_bounds) \
                else np.random.uniform(low=lo, high=hi, size=len(lo))
        elif out_of_bounds(init_hyperparameters, hyperparameter_bounds):
            warnings.warn("Your init_hyperparameters are out of bounds. They will be over-written")
            init_hyperparameters = np.random.uniform(low=lo, high=hi, size=len(lo))
        user_obj = objective_function is not None
        if method == "mcmc":
            objective_function = self.marginal_likelihood.log_likelihood
        elif objective_function is None:
            objective_function = self.marginal_likelihood.neg_log_likelihood
        if user_obj and objective_function_gradient is None and method in ("local", "hgdl"):
            raise Exception("A gradient (and Hessian) of the objective function must be provided "
                            "for method='local' or method='hgdl'.")
        if objective_function_gradient is None:
            objective_function_gradient = self.marginal_likelihood.neg_log_likelihood_gradient
        if objective_function_hessian is None:
            objective_function_hessian = self.marginal_likelihood.neg_log_likelihood_hessian
        before_hps = np.array(self.hyperparameters)
        before = self.marginal_likelihood.log_likelihood() if accept_only_if_improved and not user_obj else None
        population_objective = population_log_likelihood = None
        if not user_obj and method == "global" and self.marginal_likelihood.population_supported():
            population_objective = lambda T: -self.marginal_likelihood.log_likelihood_population(T)   # noqa: E731
        if method == "mcmc" and self.marginal_likelihood.population_supported():
            population_log_likelihood = self.marginal_likelihood.log_likelihood_population
        hps = self.trainer.train(objective_function=objective_function,
                                 population_objective=population_objective,
                                 population_log_likelihood=population_log_likelihood,
                                 objective_function_gradient=objective_function_gradient,
                                 objective_function_hessian=objective_function_hessian,
                                 hyperparameter_bounds=hyperparameter_bounds,
                                 init_hyperparameters=np.array(init_hyperparameters, dtype=np.float64), method=method,
                                 pop_size=pop_size, tolerance=tolerance, max_iter=max_iter,
                                 local_optimizer=local_optimizer, global_optimizer=global_optimizer,
                                 constraints=constraints, mcmc_prior=mcmc_prior, mcmc_prop_distrs=mcmc_prop_distrs,
                                 mcmc_args=mcmc_args, bo_args=bo_args, dask_client=dask_client, info=info)
        self.set_hyperparameters(hps)
        if before is not None and self.marginal_likelihood.log_likelihood() < before:      # gp.py:1144
            self.set_hyperparameters(before_hps)
        return self.hyperparameters

    # ---- likelihood pass-throughs (gp.py:1310-1369) ---------------------------------------------
    def log_likelihood(self, hyperparameters=None):
        return self.marginal_likelihood.log_likelihood(hyperparameters=hyperparameters)

    def neg_log_likelihood(self, hyperparameters=None):
        return self.marginal_likelihood.neg_log_likelihood(hyperparameters=hyperparameters)

    def neg_log_likelihood_gradient(self, hyperparameters=None, component=0):
        return self.marginal_likelihood.neg_log_likelihood_gradient(hyperparameters=hyperparameters,
                                                                    component=component)

    def neg_log_likelihood_hessian(self, hyperparameters=None):
        return self.marginal_likelihood.neg_log_likelihood_hessian(hyperparameters=hyperparameters)

    # ---- population entry points (beyond the reference API; what train(method="global") uses) -----
    def log_likelihood_population(self, hyperparameters):
        """LML for every row of `hyperparameters` (B, H) -> (B,): the proposals run on concurrent streams."""
        return self.marginal_likelihood.log_likelihood_population(hyperparameters)

    def neg_log_likelihood_gradient_population(self, hyperparameters, component=0):
        """grad(-LML) for every row of `hyperparameters` (B, H) -> (B, H)."""
        return self.marginal_likelihood.neg_log_likelihood_gradient_population(hyperparameters, component=component)

    def test_log_likelihood_gradient(self, hyperparameters, epsilon=1e-6):
        return self.marginal_likelihood.test_log_likelihood_gradient(hyperparameters, epsilon=epsilon)

    def log_likelihood_variance(self):
        return self.marginal_likelihood.log_likelihood_variance()

    # ---- posterior pass-throughs (gp.py:1376-1490) -----------------------------------------------
    def posterior_mean(self, x_pred, hyperparameters=None, x_out=None):
        return self.posterior.posterior_mean(x_pred, hyperparameters=hyperparameters, x_out=x_out)

    def posterior_covariance(self, x_pred, x_out=None, variance_only=False, add_noise=False):
        return self.posterior.posterior_covariance(x_pred, x_out=x_out, variance_only=variance_only,
                                                   add_noise=add_noise)

    def rmse(self, x_test, y_test):
        """gp.py:1755: root mean square error of the posterior mean."""
        return float(np.sqrt(np.mean((self.posterior_mean(x_test)["m(x)"] - y_test) ** 2)))
