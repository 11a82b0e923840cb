#!/bin/bash
# speculative MCMC through the population path: GPU suite + C1 train() wall time
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python tools/small_n_latency.py 2>&1 | tail -5
python - <<'PY'
import time, numpy as np
from fvgp_b200 import GP
rng = np.random.default_rng(1)
x = rng.random((1000, 1)); y = np.sin(5 * x[:, 0]) + np.cos(10 * x[:, 0]) + 0.05 * rng.standard_normal(1000)
for tag in ("speculative", "one at a time"):
    gp = GP(x, y, init_hyperparameters=np.array([1.0, 0.3]), noise_variances=np.full(1000, 1e-2))
    if tag != "speculative":
        gp.marginal_likelihood.population_supported = lambda want_grad=False: False
    gp.train(hyperparameter_bounds=np.array([[.01, 10], [.01, 10]]), method="mcmc", max_iter=50, mcmc_args={"seed": 0})
    t0 = time.perf_counter()
    gp.train(hyperparameter_bounds=np.array([[.01, 10], [.01, 10]]), method="mcmc", max_iter=1000, mcmc_args={"seed": 1})
    dt = time.perf_counter() - t0
    i = gp.trainer.mcmc_info
    print(f"C1 train(mcmc, 1000 updates) {tag}: {dt:.3f} s, {i['likelihood calls']} likelihood calls, acceptance {i['acceptance rate']:.2f}, hps {gp.hyperparameters}")
PY
