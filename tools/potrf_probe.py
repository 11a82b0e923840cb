"""POTRF / POTRS / LML latency at the sizes fvGP users run (N = 1000 ... 16 384), next to torch.linalg.cholesky
(cuSOLVER) on the same matrices -- the comparison bar of SURVEY 2c (ii) (gp_lin_alg.py:248-253).
FVGP_POTRF_TILE selects the diagonal-tile kernel generation (default: third)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fvgp_b200 import GP, ops  # noqa: E402
from fvgp_b200 import _lib as L  # noqa: E402


def best(fn, reps=5):
    t = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        t = min(t, a.elapsed_time(b) * 1e-3)
    return t, out


print(f"tile kernel generation: {os.environ.get('FVGP_POTRF_TILE', '3 (default)')}")
rng = np.random.default_rng(2)
for n in (1024, 2048, 4096, 8192, 16384, 50000):
    if n == 50000 and "--big" not in sys.argv:
        continue
    x = L.to_dev(rng.random((n, 3)))
    noise = L.to_dev(np.full(n, 1e-2))
    th = np.array([1.0, .3, .4, .5])
    out = L.dev_matrix(n, n)

    def fill():
        return ops.kfill(L.K_MATERN32, x, x, th[0], 1 / th[1:], 1.0, noise=noise, mode=L.FILL_LOWER, out=out)
    fill()
    ref = torch.tril(out[0][:, :n]).clone()
    t_ours = 1e30
    for _ in range(4):
        fill()
        t, f = best(lambda: ops.potrf(out[0], out[1], n), reps=1)
        t_ours = min(t_ours, t)
    Lo = torch.tril(out[0][:, :n])
    rhs = L.to_dev(rng.random((1, n)))
    t_solve, _ = best(lambda: ops.potrs(f, rhs.clone()), reps=5)
    sym = ref + torch.tril(ref, -1).T
    t_cus, Lc = best(lambda: torch.linalg.cholesky(sym), reps=3)
    err = float((Lo - Lc).abs().max() / Lc.abs().max())
    print(f"N={n}: potrf ours {t_ours * 1e3:.2f} ms ({n ** 3 / 3 / t_ours / 1e12:.1f} TFLOP/s) | torch.linalg.cholesky "
          f"{t_cus * 1e3:.2f} ms ({n ** 3 / 3 / t_cus / 1e12:.1f}) | max |L - L_cusolver| rel {err:.1e} | potrs 1 rhs "
          f"{t_solve * 1e3:.3f} ms", flush=True)
    del out, ref, sym, Lo, Lc
    torch.cuda.empty_cache()
for n, d in ((1000, 1), (4000, 3), (8000, 3)):
    rng = np.random.default_rng(1)
    xh = rng.random((n, d))
    y = np.sin(5 * xh[:, 0]) + 0.05 * rng.standard_normal(n)
    h = np.array([1.0] + [0.3] * d)
    gp = GP(xh, y, init_hyperparameters=h, noise_variances=np.full(n, 1e-2))
    for k in range(3):
        gp.log_likelihood(h * (1 + 0.01 * k))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(20):
        gp.log_likelihood(h * (1.1 + 0.01 * k))
    torch.cuda.synchronize()
    t_lml = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for k in range(10):
        gp.log_likelihood(h * (1.2 + 0.01 * k)), gp.neg_log_likelihood_gradient(h * (1.2 + 0.01 * k))
    torch.cuda.synchronize()
    print(f"N={n} D={d}: LML {t_lml * 1e3:.2f} ms, LML+gradient {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms", flush=True)
