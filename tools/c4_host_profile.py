"""Where does one gp2Scale evaluation at N = 1M go on the host side?  cProfile of GP.log_likelihood + per-step wall."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import torch  # noqa: E402

import bench  # noqa: E402
from fvgp_b200 import GP  # noqa: E402

n = 1000000
x, y, noise = bench.synthetic_c4(n)
gp = GP(x, y, init_hyperparameters=bench.theta_c4(0, n), noise_variances=noise, gp2Scale=True, linalg_mode="sparseCGpre",
        args=dict(bench.C4_ARGS))
for k in range(3):
    gp.log_likelihood(bench.theta_c4(k + 1, n))
torch.cuda.synchronize()
for k in range(5):
    t0 = time.perf_counter()
    gp.log_likelihood(bench.theta_c4(k + 4, n))
    torch.cuda.synchronize()
    print(f"step {k}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
pr = cProfile.Profile()
pr.enable()
gp.log_likelihood(bench.theta_c4(11, n))
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
for greedy in ("1", "0"):
    os.environ["FVGP_SLQ_GREEDY"] = greedy
