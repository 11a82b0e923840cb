"""Debug: which entries differ between the FULL / SYMMETRIC / LOWER fills (centred and exact-difference paths)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fvgp_b200 import _lib as L, ops

for (n, d, seed, h) in ((500, 1, 1, [1.0, 0.3]), (600, 3, 2, [1.0, .3, .4, .5]), (333, 2, 5, [1.2, .2, .7])):
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    h = np.array(h)
    xd = L.to_dev(x)
    for bounds in ((x.min(axis=0), x.max(axis=0)), None):
        outs = {}
        for name, mode in (("full", L.FILL_FULL), ("sym", L.FILL_SYMMETRIC), ("lower", L.FILL_LOWER)):
            buf, ld = L.dev_matrix(n, n)
            buf.zero_()
            ops.kfill(L.K_MATERN32, xd, xd, h[0], 1 / h[1:], 1.0, mode=mode, out=(buf, ld), bounds=bounds)
            torch.cuda.synchronize()
            outs[name] = buf[:, :n].cpu().numpy()
        for name in ("sym", "lower"):
            a, b = outs[name], outs["full"]
            if name == "lower":
                a, b = np.tril(a), np.tril(b)
            bad = np.argwhere(a != b)
            print(f"n={n} d={d} centred={bounds is not None} {name} vs full: {len(bad)} differing entries", end="")
            if len(bad):
                r, c = bad[0]
                print(f"; first {tuple(bad[0])} tile ({r // 64},{c // 64}) {a[r, c]!r} vs {b[r, c]!r}; rows {np.unique(bad[:, 0] // 64)[:10]} cols {np.unique(bad[:, 1] // 64)[:10]}"
                      f" max rel {np.max(np.abs(a - b)[a != b] / np.abs(b)[a != b]):.3e}; upper {np.sum(bad[:, 1] > bad[:, 0])} lower {np.sum(bad[:, 1] < bad[:, 0])} diag {np.sum(bad[:, 1] == bad[:, 0])}")
            else:
                print()
