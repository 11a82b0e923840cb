"""Time fill(lower) + potrf (+ potri with --potri) at size n for the current FVGP_POTRF_NB."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fvgp_b200 import _lib as L, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
rng = np.random.default_rng(2)
x = L.to_dev(rng.random((n, 3)))
noise = L.to_dev(np.full(n, 1e-2))
th = np.array([1.0, .3, .4, .5])
out = L.dev_matrix(n, n)
for rep in range(2):
    ops.kfill(L.K_MATERN32, x, x, th[0], 1 / th[1:], 1.0, noise=noise, mode=L.FILL_LOWER, out=out)
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    f = ops.potrf(out[0], out[1], n)
    b.record()
    if "--potri" in sys.argv:
        ops.potri(f)
    c.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3
    msg = f"nb={os.environ.get('FVGP_POTRF_NB', 'default')} n={n} potrf {t * 1e3:.1f} ms = {n ** 3 / 3 / t / 1e12:.2f} TFLOP/s"
    if "--potri" in sys.argv:
        t2 = b.elapsed_time(c) * 1e-3
        msg += f"; potri {t2 * 1e3:.1f} ms = {2 * n ** 3 / 3 / t2 / 1e12:.2f} TFLOP/s"
    print(msg, flush=True)
