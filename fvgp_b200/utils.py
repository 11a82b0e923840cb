"""Host-side helpers."""
import numpy as np


def morton_order(x, bits=10):
    """Permutation that sorts points along a Z-order (Morton) curve.

    gp2Scale's block / tile culls are effective when consecutive points are spatial neighbours;
    the reference recommends a locality-preserving order for the same reason (kernels.py:607-616).
    The permutation is applied to x AND y before constructing the GP; it changes the layout of the
    covariance matrix, not its entries."""
    x = np.asarray(x, dtype=np.float64)
    lo, hi = x.min(axis=0), x.max(axis=0)
    q = np.minimum(((x - lo) / np.where(hi > lo, hi - lo, 1.0) * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    d = x.shape[1]
    key = np.zeros(len(x), dtype=np.uint64)
    for b in range(bits):
        for i in range(d):
            key |= ((q[:, i] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * d + i)
    return np.argsort(key, kind="stable")
