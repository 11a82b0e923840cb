"""Latency of one LML / LML+gradient through the public API at small N (C1 scale)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fvgp_b200 import GP, _lib as L

for n, d in ((1000, 1), (2000, 3), (4000, 3), (8000, 3)):
    rng = np.random.default_rng(1)
    x = rng.random((n, d))
    y = np.sin(5 * x[:, 0]) + 0.05 * rng.standard_normal(n)
    h = np.array([1.0] + [0.3] * d)
    gp = GP(x, y, init_hyperparameters=h, noise_variances=np.full(n, 1e-2))
    lib = L.load()
    for k in range(3):
        gp.log_likelihood(h * (1 + 0.01 * k))
    torch.cuda.synchronize()
    l0 = lib.fvgp_launch_count()
    t0 = time.perf_counter()
    reps = 20
    for k in range(reps):
        gp.log_likelihood(h * (1.1 + 0.01 * k))
    torch.cuda.synchronize()
    t_lml = (time.perf_counter() - t0) / reps
    launches = (lib.fvgp_launch_count() - l0) / reps
    t0 = time.perf_counter()
    for k in range(reps):
        gp.log_likelihood(h * (1.2 + 0.01 * k)); gp.neg_log_likelihood_gradient(h * (1.2 + 0.01 * k))
    torch.cuda.synchronize()
    t_both = (time.perf_counter() - t0) / reps
    print(f"N={n} D={d}: LML {t_lml * 1e3:.2f} ms ({launches:.0f} launches), LML+gradient {t_both * 1e3:.2f} ms", flush=True)
t0 = time.perf_counter()
rng = np.random.default_rng(1)
x = rng.random((1000, 1)); y = np.sin(5 * x[:, 0]) + np.cos(10 * x[:, 0]) + 0.05 * rng.standard_normal(1000)
gp = GP(x, y, init_hyperparameters=np.array([1.0, 0.3]), noise_variances=np.full(1000, 1e-2))
t0 = time.perf_counter()
gp.train(hyperparameter_bounds=np.array([[.01, 10], [.01, 10]]), method="mcmc", max_iter=200)
print(f"C1 train(method='mcmc', max_iter=200): {time.perf_counter() - t0:.2f} s -> hps {gp.hyperparameters}")
