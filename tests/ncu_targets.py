"""Small single-purpose workloads for `ncu` captures (see profiles/README.md).

    python tests/ncu_targets.py gemm|syrk|kfill|kfill_lower|trace|wendland|spmv|slq|pcg [n]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fvgp_b200 import _lib as L  # noqa: E402
from fvgp_b200 import ops  # noqa: E402

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
rng = np.random.default_rng(0)
reps = 3
if what in ("gemm", "syrk"):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        ops.dgemm_nt(a, a, c, alpha=-1.0, beta=1.0, lower=(what == "syrk"))
elif what in ("kfill", "kfill_lower", "kfill_full"):
    x = L.to_dev(rng.random((n, 3)))
    noise = L.to_dev(np.full(n, 1e-2))
    out = L.dev_matrix(n, n)
    bounds = None if os.environ.get("FVGP_EXACT_DIFF") else (np.zeros(3), np.ones(3))
    mode = {"kfill": L.FILL_SYMMETRIC, "kfill_lower": L.FILL_LOWER, "kfill_full": L.FILL_FULL}[what]
    for _ in range(reps):
        ops.kfill(L.K_MATERN32, x, x, 1.0, 1 / np.array([.3, .4, .5]), 1.0, noise=noise, mode=mode, out=out,
                  bounds=bounds)
elif what == "trace":
    x = L.to_dev(rng.random((n, 3)))
    buf, ld = L.dev_matrix(n, n)
    buf.normal_()
    b = L.to_dev(rng.standard_normal(n))
    for _ in range(reps):
        ops.kgrad_trace_matern32(x, np.array([1.0, .3, .4, .5]), buf, ld, b)
elif what in ("wendland", "spmv", "slq", "pcg"):
    xs = rng.random((n, 3))
    xs = L.to_dev(xs[np.lexsort((xs[:, 2] // .05, xs[:, 1] // .05, xs[:, 0] // .05))])
    th = np.array([1.0, .029, .029, .029])
    th[1:] *= (1e6 / n) ** (1 / 3)
    for _ in range(reps if what == "wendland" else 1):
        K = ops.wendland_csr(xs, xs, th)
    if what == "slq":
        ops.slq_logdet(K, degree=3, probes=8, seed=0)
    if what == "pcg":
        ops.pcg(K, L.to_dev(rng.standard_normal(n)), rtol=1e-30, maxiter=3)
    if what == "spmv":
        v = L.to_dev(rng.standard_normal(n))
        y = L.dev_empty((n,))
        for _ in range(reps):
            ops.spmv(K, v, y)
torch.cuda.synchronize()
print("done", what, n)
