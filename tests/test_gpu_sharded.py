"""GPU parity of the block-cyclic dense path (fvgp_b200/sharded.py) through the C ABI.

world = 1 runs the whole choreography (blocked potrf / trsm / updates / solves / TRTRI / LAUUM / block traces)
on one GPU against the oracle; the 2-rank test needs two GPUs (skipped otherwise) and uses NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _problem(n, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3))
    y = np.sin(5 * x[:, 0]) * np.cos(3 * x[:, 1]) + x[:, 2] + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2) + 1e-3 * rng.random(n), np.array([1.1, .3, .45, .5])


@pytest.mark.parametrize("n,nb", [(700, 128), (1500, 256), (2100, 512), (1024, 256), (90, 128)])
def test_sharded_single_rank_matches_oracle(n, nb):
    from fvgp_b200 import sharded
    from oracle import fvgp_oracle as orc
    x, y, noise, theta = _problem(n)
    ev = sharded.ShardedDenseEvaluator(x, y, noise, nb=nb, grid=(1, 1))
    out = ev.evaluate(0, theta[0], 1.0 / theta[1:], 1.0, np.full(n, y.mean()), want_gradient_theta=theta)
    lml_ref = orc.dense_log_likelihood(x, y, theta, noise)
    grad_ref = orc.dense_neg_log_likelihood_gradient(x, y, theta, noise, economical=True)
    assert abs(out["lml"] / lml_ref - 1) < 1e-8
    assert np.max(np.abs(0.5 * out["traces"] - grad_ref) / np.abs(grad_ref)) < 1e-8


def test_sharded_reports_non_positive_definite():
    from fvgp_b200 import _lib as L
    from fvgp_b200 import sharded
    x, y, noise, theta = _problem(600)
    x[300] = x[10]                                    # duplicate point, no noise -> singular
    ev = sharded.ShardedDenseEvaluator(x, y, np.zeros(600), nb=128, grid=(1, 1))
    with pytest.raises(L.NonPositiveDefiniteError):
        ev.evaluate(0, theta[0], 1.0 / theta[1:], 1.0, np.full(600, y.mean()))


def _worker(rank, world, port, n, nb, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group(backend="nccl", rank=rank, world_size=world)
    from fvgp_b200 import sharded
    x, y, noise, theta = _problem(n)
    ev = sharded.ShardedDenseEvaluator(x, y, noise, nb=nb)
    out = ev.evaluate(0, theta[0], 1.0 / theta[1:], 1.0, np.full(n, y.mean()), want_gradient_theta=theta)
    torch.cuda.synchronize()
    q.put((rank, out["lml"], out["traces"], ev.comm.bytes_received))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_two_ranks_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from oracle import fvgp_oracle as orc
    n, nb = 3000, 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, nb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x, y, noise, theta = _problem(n)
    lml_ref = orc.dense_log_likelihood(x, y, theta, noise)
    grad_ref = orc.dense_neg_log_likelihood_gradient(x, y, theta, noise, economical=True)
    for rank, lml, traces, recv in res:
        assert abs(lml / lml_ref - 1) < 1e-8
        assert np.max(np.abs(0.5 * traces - grad_ref) / np.abs(grad_ref)) < 1e-8
        assert recv > 0
