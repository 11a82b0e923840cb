"""Timing harness of the block-cyclic dense path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        tests/sharded_bench.py --size 60000 [--nb 2048] [--grad] [--check]

Prints one JSON line on rank 0: seconds per phase (CUDA events, max over ranks), TFLOP/s per GPU on the
algorithmic N^3/3 (+ 2N^3/3 with --grad), bytes received per rank."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fvgp_b200 import parallel, sharded  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", dest="n", type=int, default=40000,
                help="matrix order N (not --n: torchrun's own parser treats a bare --n as an ambiguous abbreviation)")
ap.add_argument("--nb", type=int, default=0)
ap.add_argument("--dim", type=int, default=3)
ap.add_argument("--grad", action="store_true")
ap.add_argument("--check", action="store_true", help="compare with the single-GPU path on rank 0 (n must fit)")
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()

rank, local_rank, world = parallel.init()
rng = np.random.default_rng(5)
n = args.n
x = rng.random((n, args.dim))
y = np.sin(5 * x[:, 0]) + 0.1 * rng.standard_normal(n)
noise = np.full(n, 1e-2)
theta = np.array([1.0] + [.3, .4, .5, .35][:args.dim])
ev = sharded.ShardedDenseEvaluator(x, y, noise, nb=args.nb or None)
A = ev._matrix()
mean = np.full(n, y.mean())


def timed(fn):
    torch.cuda.synchronize()
    parallel.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    return out, parallel.max_over_ranks(a.elapsed_time(b) * 1e-3)


res = {}
for rep in range(args.reps):
    _, t_fill = timed(lambda: A.fill(0, ev.x_dev, ev.x, theta[0], 1 / theta[1:], 1.0, ev.noise_dev,
                                     (ev.bounds[0] + ev.bounds[1]) / 2))
    info, t_fac = timed(A.factor)
    rhs = ev.ops.upload(y - mean)
    alpha, t_solve = timed(lambda: A.solve(rhs))
    logdet, t_ld = timed(A.logdet)
    lml = float(-0.5 * (float((rhs * alpha).sum().item()) + logdet + n * np.log(2 * np.pi)))
    res = {"n": n, "nb": ev.nb, "grid": list(ev.grid), "world": world, "info": info, "lml": lml,
           "t_fill": t_fill, "t_factor": t_fac, "t_solve": t_solve, "t_logdet": t_ld,
           "factor_tflops_per_gpu": n ** 3 / 3 / t_fac / 1e12 / world,
           "local_GB": A.local_bytes() / 1e9, "recv_GB_factor": ev.comm.bytes_received / 1e9}
    if args.grad:
        r0 = ev.comm.bytes_received
        _, t_inv = timed(A.invert)
        tr, t_tr = timed(lambda: A.grad_traces(theta, alpha))
        res.update({"t_invert": t_inv, "t_traces": t_tr, "invert_tflops_per_gpu": 2 * n ** 3 / 3 / t_inv / 1e12 / world,
                    "recv_GB_invert": (ev.comm.bytes_received - r0) / 1e9, "grad": (0.5 * tr).tolist()})
    res["t_total"] = sum(v for k, v in res.items() if k.startswith("t_"))
    ev.comm.bytes_received = 0
if args.check and rank == 0:
    from fvgp_b200 import GP
    gp = GP(x, y, init_hyperparameters=theta, noise_variances=noise)
    t0 = time.time()
    ref = gp.log_likelihood(theta)
    res["single_gpu_lml"] = ref
    res["lml_rel_diff"] = abs(res["lml"] / ref - 1)
    if args.grad:
        g = gp.neg_log_likelihood_gradient(theta)
        res["grad_rel_diff"] = float(np.max(np.abs(np.array(res["grad"]) - g) / np.abs(g)))
    res["single_gpu_seconds"] = time.time() - t0
if rank == 0:
    print(json.dumps(res))
parallel.barrier()
if dist.is_initialized():
    dist.destroy_process_group()
