"""POTRF at N = 50 000 for the block width in FVGP_POTRF_NB (one process per width: the setting is read once)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fvgp_b200 import ops  # noqa: E402
from fvgp_b200 import _lib as L  # noqa: E402

n = int(os.environ.get("PROBE_N", "50000"))
rng = np.random.default_rng(2)
x = L.to_dev(rng.random((n, 3)))
noise = L.to_dev(np.full(n, 1e-2))
th = np.array([1.0, .3, .4, .5])
out = L.dev_matrix(n, n)
times = []
for _ in range(4):
    ops.kfill(L.K_MATERN32, x, x, th[0], 1 / th[1:], 1.0, noise=noise, mode=L.FILL_LOWER, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.potrf(out[0], out[1], n)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b) * 1e-3)
print(f"N={n} nb={os.environ.get('FVGP_POTRF_NB', 'default')} ozaki={os.environ.get('FVGP_OZAKI', 'default')}: "
      f"potrf {min(times):.3f} s (all: {' '.join(f'{t:.3f}' for t in times)}) -> {n ** 3 / 3 / min(times) / 1e12:.1f} TFLOP/s", flush=True)
