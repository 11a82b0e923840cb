"""update_gp_data(append=True): bordered Cholesky update vs a full refactorisation."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fvgp_b200 import GP

for n, m in ((30000, 10), (30000, 500), (50000, 10)):
    rng = np.random.default_rng(3)
    x = rng.random((n + m, 3))
    y = np.sin(5 * x[:, 0]) * np.cos(3 * x[:, 1]) + x[:, 2] + 0.1 * rng.standard_normal(n + m)
    nz = np.full(n + m, 1e-2)
    h = np.array([1.0, .3, .4, .5])
    gp = GP(x[:n], y[:n], init_hyperparameters=h, noise_variances=nz[:n])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gp.update_gp_data(x[n:], y[n:], noise_variances_new=nz[n:], append=True)
    torch.cuda.synchronize()
    t_app = time.perf_counter() - t0
    lml_app = gp.log_likelihood()
    t0 = time.perf_counter()
    gp.set_hyperparameters(h)                    # full refill + refactorisation of the (n + m)-point GP
    torch.cuda.synchronize()
    t_full = time.perf_counter() - t0
    print(f"N={n} + {m} appended: bordered update {t_app * 1e3:.1f} ms (appended_rows={m}), full refresh {t_full * 1e3:.1f} ms, "
          f"LML rel diff {abs(lml_app / gp.log_likelihood() - 1):.2e}", flush=True)
    del gp
