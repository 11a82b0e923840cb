"""Time SpMV / PCG / SLQ on the C4 matrix (N = 1M by default) for the current FVGP_SPMV_VARIANT."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fvgp_b200 import _lib as L, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
x, y, noise = bench.synthetic_c4(n)
th = bench.theta_c4(1, n)
xd, nd = L.to_dev(x), L.to_dev(noise)
KV = ops.wendland_csr(xd, xd, th, noise=nd)
v = L.to_dev(y - y.mean())
out = L.dev_empty((n,))


def timed(fn, reps):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best, r


t, _ = timed(lambda: ops.spmv(KV, v, out), 20)
ref = KV.to_scipy() @ (y - y.mean()) if n <= 200000 else None
tag = f"variant={os.environ.get('FVGP_SPMV_VARIANT', 'default')} n={n} nnz={KV.nnz}"
print(f"{tag} spmv {t * 1e6:.0f} us -> {(12.0 * KV.nnz + 16.0 * n) / t / 1e9:.0f} GB/s", flush=True)
if ref is not None:
    print("   max rel err vs scipy", float(np.max(np.abs(out.cpu().numpy() - ref)) / np.max(np.abs(ref))))
t, res = timed(lambda: ops.pcg(KV, v, rtol=1e-5), 2)
print(f"{tag} pcg plain {t * 1e3:.1f} ms, {res[2]} iterations -> {t / max(res[2], 1) * 1e6:.0f} us/iteration", flush=True)
M = ops.bjacobi(KV)
t, res = timed(lambda: ops.pcg(KV, v, rtol=1e-5, precond=M), 2)
print(f"{tag} pcg bjacobi {t * 1e3:.1f} ms, {res[2]} iterations -> {t / max(res[2], 1) * 1e6:.0f} us/iteration", flush=True)
t, _ = timed(lambda: ops.slq_logdet(KV, degree=20, probes=10, seed=0), 2)
print(f"{tag} slq 10x20 {t * 1e3:.1f} ms", flush=True)
t, _ = timed(lambda: ops.slq_logdet(KV, degree=20, probes=16, seed=0), 2)
print(f"{tag} slq 16x20 {t * 1e3:.1f} ms", flush=True)
