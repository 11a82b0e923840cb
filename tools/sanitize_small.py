"""Tiny workload for compute-sanitizer (racecheck / memcheck) over the kernels added in this round: blocked tile
Cholesky, fused substitutions, batched fill, lock-step population (BATCHED GEMM), gradient traces."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvgp_b200 import GP
rng = np.random.default_rng(0)
n = 300
x = rng.random((n, 2)); y = np.sin(4 * x[:, 0]) + x[:, 1] + 0.1 * rng.standard_normal(n)
h = np.array([1.0, 0.3, 0.4])
gp = GP(x, y, init_hyperparameters=h, noise_variances=np.full(n, 1e-2))
a = gp.log_likelihood(h * 1.1)
g = gp.neg_log_likelihood_gradient(h * 1.1)
T = h * np.array([[1.1, 1.1, 1.1], [0.9, 1.2, 1.0], [1.3, 0.8, 0.7]])
lml, grad = gp.marginal_likelihood.evaluate_population(T, with_gradient=True)
print("ok", a == lml[0], np.max(np.abs(grad[0] - g) / np.abs(g)))
