"""Training drivers (reference: fvgp/gp_training.py, gp_mcmc.py).  Pure callers of the hot path:
they receive `log_likelihood` / `neg_log_likelihood(_gradient)` callables and are host control
loops (out of the kernel scope).  global / local use the same scipy
optimisers as the reference; mcmc is a compact adaptive Metropolis-Hastings with the reference's
contract (uniform prior on the bounds, result = median of the chain, gp_training.py:146-162);
hgdl / bo / asynchronous training need third-party packages this build does not bundle."""
import warnings

import numpy as np
from scipy.optimize import differential_evolution, minimize


class GPtraining:
    def __init__(self, data, hyperparameters):
        self.data = data
        self.hyperparameters = np.array(hyperparameters, dtype=np.float64)
        self.mcmc_info = None

    @staticmethod
    def _in_bounds(v, bounds):
        return not (np.any(v < bounds[:, 0]) or np.any(v > bounds[:, 1]))

    def train(self, objective_function=None, objective_function_gradient=None, objective_function_hessian=None,
              hyperparameter_bounds=None, init_hyperparameters=None, method="global", pop_size=20, tolerance=0.0001,
              max_iter=120, local_optimizer="L-BFGS-B", global_optimizer="genetic", constraints=(), mcmc_prior=None,
              mcmc_prop_distrs="normal", mcmc_args={}, bo_args=None, dask_client=None, info=False,
              population_objective=None, population_log_likelihood=None):
        if not self._in_bounds(init_hyperparameters, hyperparameter_bounds):
            raise Exception("Starting positions outside of optimization bounds.", init_hyperparameters,
                            hyperparameter_bounds)
        if method == "global" and population_objective is not None:
            # A whole generation per call: scipy hands the trial population over as (H, S) and the S proposals run
            # on concurrent streams (GPMarginalLikelihood.evaluate_population).  `vectorized` implies deferred
            # updating (the best member is refreshed once per generation, not after every individual).
            res = differential_evolution(lambda X: population_objective(np.atleast_2d(X.T)), hyperparameter_bounds,
                                         maxiter=max_iter, popsize=pop_size, tol=tolerance, disp=info, polish=False,
                                         x0=init_hyperparameters.reshape(1, -1)[0], constraints=constraints,
                                         vectorized=True, updating="deferred")
            return np.array(res["x"])
        if method == "global":
            res = differential_evolution(objective_function, hyperparameter_bounds, maxiter=max_iter, popsize=pop_size,
                                         tol=tolerance, disp=info, polish=False,
                                         x0=init_hyperparameters.reshape(1, -1)[0], constraints=constraints, workers=1)
            return np.array(res["x"])
        if method == "local":
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                uses_hess = local_optimizer in ("Newton-CG", "dogleg", "trust-ncg", "trust-krylov", "trust-exact",
                                                "trust-constr")
                res = minimize(objective_function, init_hyperparameters, method=local_optimizer,
                               jac=objective_function_gradient,
                               hess=objective_function_hessian if uses_hess else None,
                               bounds=hyperparameter_bounds, tol=tolerance, constraints=constraints,
                               options={"maxiter": max_iter})
            return np.array(res["x"])
        if method == "mcmc":
            res = run_mcmc(objective_function, init_hyperparameters, hyperparameter_bounds, n_updates=max_iter,
                           prior=mcmc_prior, info=info, seed=mcmc_args.get("seed", None) if mcmc_args else None,
                           population_log_likelihood=population_log_likelihood)
            self.mcmc_info = res
            return res["median(x)"]
        if method == "adam":
            return adam_optimize(objective_function, objective_function_gradient, init_hyperparameters,
                                 hyperparameter_bounds, max_iter=max_iter, tolerance=tolerance)
        if callable(method):
            return np.asarray(method(self))
        raise Exception(f"method '{method}' is not available in fvgp_b200 (supported: global, local, mcmc, adam, "
                        "or a callable); hgdl / bo need the hgdl / gp_bo packages")


def run_mcmc(log_likelihood, x0, bounds, n_updates=10000, prior=None, info=False, seed=None,
             population_log_likelihood=None, max_batch=16):
    """Adaptive random-walk Metropolis-Hastings over the hyperparameters (cf. gp_mcmc.py:96-224).

    population_log_likelihood(T (B, H)) -> (B,): speculative evaluation.  While the chain rejects it stays where it
    is, so the next proposals are all drawn around the SAME state: a batch of them is evaluated in one population
    call and consumed in order up to (and including) the first acceptance; the rest -- drawn around a state the
    chain has left, never looked at -- is discarded.  Every step still uses fresh independent randomness and the
    exact acceptance ratio, so this is the same Markov chain; with acceptance rate a it advances ~1/a steps per call."""
    rng = np.random.default_rng(seed)
    x = np.array(x0, dtype=float)
    span = bounds[:, 1] - bounds[:, 0]
    step = 0.05 * span

    def log_prior(t):
        if prior is not None:
            return prior(t, bounds, {})
        return 0.0 if GPtraining._in_bounds(t, bounds) else -np.inf

    f = log_likelihood(x) + log_prior(x)
    chain, fs, accepted = [x.copy()], [f], 0
    n_updates = int(n_updates)
    it = 0
    calls = 0
    while it < n_updates:
        # batch: never across an adaptation boundary (the step width changes there), sized to the acceptance rate
        rate = accepted / it if it >= 20 else 0.3
        want = 1 if population_log_likelihood is None else int(np.clip(np.ceil(1.5 / max(rate, 0.05)), 2, max_batch))
        m = max(1, min(want, n_updates - it, 50 - it % 50))
        props = x + step * rng.standard_normal((m, len(x)))
        us = rng.random(m)
        lps = np.array([log_prior(p) for p in props], dtype=float)
        ok = np.isfinite(lps)
        fps = np.full(m, -np.inf)
        lazy = population_log_likelihood is None or ok.sum() <= 1
        if not lazy:
            try:
                fps[ok] = np.asarray(population_log_likelihood(props[ok]), dtype=float) + lps[ok]
                calls += 1
            except Exception:
                lazy = True           # e.g. a speculative proposal that is not positive definite: decide one by one
        for i in range(m):
            took = False
            if ok[i]:
                if lazy:
                    fps[i] = log_likelihood(props[i]) + lps[i]
                    calls += 1
                if np.isnan(fps[i]):
                    raise Exception("NaN log-likelihood encountered in MCMC")
                if np.log(us[i]) < fps[i] - f:
                    x, f, accepted, took = props[i].copy(), fps[i], accepted + 1, True
            chain.append(x.copy())
            fs.append(f)
            it += 1
            if it % 50 == 0:                                    # adapt towards ~25-45 % acceptance
                r = accepted / it
                step *= 1.25 if r > 0.45 else (0.8 if r < 0.2 else 1.0)
            if info and it % 100 == 0:
                print(f"mcmc iteration {it}: f(x)= {f}")
            if took:
                break                                           # the remaining proposals were drawn around the old state
    chain = np.array(chain)
    burn = chain[len(chain) // 5:]
    best = int(np.argmax(fs))
    return {"x": chain, "f(x)": np.array(fs), "median(x)": np.median(burn, axis=0), "mean(x)": np.mean(burn, axis=0),
            "var(x)": np.var(burn, axis=0), "max x": chain[best], "max f(x)": fs[best],
            "acceptance rate": accepted / max(1, n_updates), "likelihood calls": calls}


def adam_optimize(objective, gradient, x0, bounds, max_iter=200, tolerance=1e-4, lr=0.02, b1=0.9, b2=0.999):
    """Projected Adam on the negative log-likelihood (cf. gp_training.py:577-667)."""
    x = np.array(x0, dtype=float)
    m, v = np.zeros_like(x), np.zeros_like(x)
    scale = bounds[:, 1] - bounds[:, 0]
    for t in range(1, int(max_iter) + 1):
        g = gradient(x)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        stepv = lr * scale * (m / (1 - b1 ** t)) / (np.sqrt(v / (1 - b2 ** t)) + 1e-12)
        x_new = np.clip(x - stepv, bounds[:, 0], bounds[:, 1])
        if np.linalg.norm(x_new - x) < tolerance * np.linalg.norm(scale) * 1e-2:
            x = x_new
            break
        x = x_new
    return x
