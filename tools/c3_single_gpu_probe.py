"""C3 (fvGP 20 000 points x 5 tasks -> N = 100 000) LML + gradient on ONE GPU in the default arithmetic: does the scratch
ladder of the INT8 POTRI find room next to the 80 GB matrix, and what does the step take."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import torch  # noqa: E402

import bench  # noqa: E402
from fvgp_b200 import fvGP, ops  # noqa: E402

pts = int(os.environ.get("PROBE_POINTS", "20000"))
x, y, noise = bench.synthetic_c3(pts)
torch.cuda.empty_cache()
gp = fvGP(x, y, init_hyperparameters=bench.THETA_C3, noise_variances=noise, args={"dense_sharded": False})
for k in range(int(os.environ.get('PROBE_STEPS', '1'))):
    th = bench.THETA_C3 * (1.01 + 0.01 * k)            # not the constructor's theta: a full fill + POTRF + POTRI
    ops.start_phase_timing()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lml = gp.log_likelihood(th)
    grad = gp.neg_log_likelihood_gradient(th)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ph = ops.stop_phase_timing()
    print(f"N={pts * y.shape[1]} step {k}: {dt:.2f} s (potrf {ph.get('potrf', 0):.2f}, potri {ph.get('potri', 0):.2f}); LML {lml:.6f}; "
          f"free HBM {torch.cuda.mem_get_info()[0] / 1e9:.0f} GB", flush=True)
