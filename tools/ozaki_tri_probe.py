"""POTRI at N = 50 000 with the triangular products on DMMA vs as chunked INT8-slice GEMMs (fvgp_set_ozaki_tri):
per-step POTRF / POTRI time (CUDA events) and agreement of LML / gradient with the all-DMMA path."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import torch  # noqa: E402

import bench  # noqa: E402
from fvgp_b200 import GP, ops  # noqa: E402
from fvgp_b200 import _lib as L  # noqa: E402

lib = L.load()
n = int(os.environ.get("PROBE_N", "50000"))
x, y, noise = bench.synthetic_c2(n)
gp = GP(x, y, init_hyperparameters=bench.theta_k(0), noise_variances=noise)
res = {}
modes = [(0, 0)] + [(8, int(t)) for t in os.environ.get('PROBE_TRI', '0,4,8,0,4,8').split(',')]
min_rows = int(os.environ.get('PROBE_MIN_ROWS', '8192'))
assert lib.fvgp_set_ozaki_gate(40000, min_rows) == 0
print('smallest leading block on the INT8 path inside POTRI:', min_rows, 'rows')
for oz, tri in modes:
    lib.fvgp_set_ozaki(oz)
    lib.fvgp_set_ozaki_tri(tri)
    for k in range(2):
        th = bench.theta_k(k + 1)
        gp.kv._memo = None
        ops.start_phase_timing()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lml = gp.log_likelihood(th)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        grad = gp.neg_log_likelihood_gradient(th)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ph = ops.stop_phase_timing()
        print(f"ozaki={oz} tri={tri} step {k}: LML {t1 - t0:.3f} s (potrf {ph.get('potrf', 0):.3f}), gradient {t2 - t1:.3f} s "
              f"(potri {ph.get('potri', 0):.3f}); free HBM {torch.cuda.mem_get_info()[0] / 1e9:.0f} GB", flush=True)
        res[(oz, tri, k)] = (lml, grad)
for oz, tri in sorted(set(modes[1:])):
    for k in range(2):
        a, b = res[(0, 0, k)], res[(oz, tri, k)]
        print(f"ozaki={oz} tri={tri} theta {k}: LML rel diff {abs(a[0] / b[0] - 1):.2e}, "
              f"gradient max rel diff {np.max(np.abs(a[1] - b[1]) / np.abs(a[1])):.2e}")
