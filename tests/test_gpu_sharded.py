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


def test_gp_api_dense_sharded_single_rank(golden):
    """args["dense_sharded"] routes GP.log_likelihood / gradient / posterior through the block-cyclic evaluator."""
    from fvgp_b200 import GP
    g = golden("dense_lml_c2")
    x, y, nz = g["x"], g["y"], g["noise"]
    gp = GP(x, y, init_hyperparameters=g["h0"], noise_variances=nz, args={"dense_sharded": True, "dense_sharded_block": 128})
    assert gp.kv.state.sharded is not None
    assert abs(gp.log_likelihood() / g["lml_h0_state"] - 1) <= 1e-8
    for hk in ("h0", "h1"):
        assert abs(gp.log_likelihood(g[hk]) / g["lml_" + hk] - 1) <= 1e-8
        grad = gp.neg_log_likelihood_gradient(g[hk])
        assert np.max(np.abs(grad - g["grad_" + hk]) / np.abs(g["grad_" + hk])) <= 1e-8
        assert abs(gp.log_likelihood(g[hk]) / g["lml_" + hk] - 1) <= 1e-8      # memo after the in-place inverse
    pm = gp.posterior_mean(g["x_pred"])                                          # state rebuilt after the matrix moved on
    assert np.allclose(pm["m(x)"], g["post_mean"], rtol=1e-8, atol=1e-10)
    pc = gp.posterior_covariance(g["x_pred"])
    assert np.allclose(pc["v(x)"], g["post_var"], rtol=1e-6, atol=1e-9)


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
    # the same through the public API: every rank holds a replica of the (small) host data and calls collectively
    from fvgp_b200 import GP
    gp = GP(x, y, init_hyperparameters=theta, noise_variances=noise, args={"dense_sharded": True, "dense_sharded_block": nb})
    api_lml = gp.log_likelihood(theta * 1.01)
    api_grad = gp.neg_log_likelihood_gradient(theta * 1.01)
    q.put((rank, out["lml"], out["traces"], ev.comm.bytes_received, api_lml, api_grad))
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
    api_lml_ref = orc.dense_log_likelihood(x, y, theta * 1.01, noise)
    api_grad_ref = orc.dense_neg_log_likelihood_gradient(x, y, theta * 1.01, noise, economical=True)
    for rank, lml, traces, recv, api_lml, api_grad in res:
        assert abs(lml / lml_ref - 1) < 1e-8
        assert np.max(np.abs(0.5 * traces - grad_ref) / np.abs(grad_ref)) < 1e-8
        assert recv > 0
        assert abs(api_lml / api_lml_ref - 1) < 1e-8
        assert np.max(np.abs(api_grad - api_grad_ref) / np.abs(api_grad_ref)) < 1e-8
