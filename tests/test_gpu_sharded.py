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


def test_sharded_single_rank_int8_slice_updates_match_oracle():
    """The block-cyclic path sends updates of >= 4096 rows through the INT8-slice GEMM (fvgp_ozaki_gemm: all three
    operand layouts of factor / invert); here the thresholds are lowered so that the same routing is exercised at a
    size the oracle covers.  n is not a multiple of the block: the ragged last block (k % 16 != 0) must fall back."""
    from fvgp_b200 import _lib as L
    from fvgp_b200 import sharded
    from oracle import fvgp_oracle as orc
    if not L.load().fvgp_ozaki_slices():
        pytest.skip("INT8-slice path not available / switched off")
    n, nb = 4200, 512
    x, y, noise, theta = _problem(n)
    res = {}
    for int8 in (False, True):
        ops = sharded.CudaLocalOps()
        ops.int8_min_rows, ops.int8_min_block = (1024, 256) if int8 else (1 << 30, 1 << 30)
        ev = sharded.ShardedDenseEvaluator(x, y, noise, nb=nb, grid=(1, 1), ops=ops)
        res[int8] = ev.evaluate(0, theta[0], 1.0 / theta[1:], 1.0, np.full(n, y.mean()), want_gradient_theta=theta)
        assert (getattr(ops, "int8_calls", 0) > 20) == int8
    lml_ref = orc.dense_log_likelihood(x, y, theta, noise)
    grad_ref = orc.dense_neg_log_likelihood_gradient(x, y, theta, noise, economical=True)
    for out in res.values():
        assert abs(out["lml"] / lml_ref - 1) < 1e-8
        assert np.max(np.abs(0.5 * out["traces"] - grad_ref) / np.abs(grad_ref)) < 1e-8
    assert not np.array_equal(res[True]["traces"], res[False]["traces"])


def test_sharded_reports_non_positive_definite():
    from fvgp_b200 import _lib as L
    from fvgp_b200 import sharded
    x, y, noise, theta = _problem(600)
    x[300] = x[10]                                    # duplicate point ...
    bad = np.zeros(600)
    bad[300] = -1e-6                                  # ... and a slightly negative "noise" there: pivot 301 is < 0 whatever
    ev = sharded.ShardedDenseEvaluator(x, y, bad, nb=128, grid=(1, 1))      # the rounding of the exact-zero Schur complement
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
    ev.ops.int8_min_rows, ev.ops.int8_min_block = 512, 128      # INT8-slice updates at this size too (default: >= 4096 rows)
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


def test_gp_api_dense_sharded_user_kernel_gradient():
    """Sharded dense gradient of user kernels composed of the fvgp.kernels names (radial block traces + descriptor
    Jacobian) equals the single-GPU fused gradient, and both match central differences of the LML."""
    from fvgp_b200 import GP
    from fvgp_b200 import kernels as K
    x, y, noise, _ = _problem(900)

    def m52(x1, x2, h):
        return h[0] * K.matern_kernel_diff2(K.get_anisotropic_distance_matrix(x1, x2, h[1:4]), 1.0)

    def se(x1, x2, h):
        return h[0] * K.squared_exponential_kernel(K.get_distance_matrix(x1, x2), h[1])
    for kern, h in ((m52, np.array([1.2, .4, .5, .6])), (se, np.array([0.9, .35]))):
        one = GP(x, y, init_hyperparameters=h, noise_variances=noise, kernel_function=kern)
        sh = GP(x, y, init_hyperparameters=h, noise_variances=noise, kernel_function=kern,
                args={"dense_sharded": True, "dense_sharded_block": 256})
        assert sh.kv.state.sharded is not None
        h1 = h * 1.03
        g1, g2 = one.neg_log_likelihood_gradient(h1), sh.neg_log_likelihood_gradient(h1)
        assert abs(one.log_likelihood(h1) / sh.log_likelihood(h1) - 1) <= 1e-10
        assert np.max(np.abs(g1 - g2) / np.abs(g1)) <= 1e-8, (g1, g2)
        fd = np.array([(one.log_likelihood(h1 + e) - one.log_likelihood(h1 - e)) / 2e-6
                       for e in 1e-6 * np.eye(len(h))])
        assert np.max(np.abs(-g2 - fd) / np.abs(fd)) <= 1e-4


def _sparse_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group(backend="nccl", rank=rank, world_size=world)
    from fvgp_b200 import GP, ops
    from fvgp_b200 import _lib as L
    from fvgp_b200.utils import morton_order
    rng = np.random.default_rng(11)
    x = rng.random((n, 3))
    x[:n // 5] = 0.3 + 0.05 * rng.standard_normal((n // 5, 3))            # a cluster: equal rows are not equal work
    x = np.ascontiguousarray(x[morton_order(x)])
    y = np.sin(6 * x[:, 0]) + 0.1 * rng.standard_normal(n)
    noise = np.full(n, 1e-2)
    th = np.array([1.1, .06, .07, .065])
    args = {"sparse_cg_tol": 1e-10, "random_logdet_min_num_samples": 12, "random_logdet_max_num_samples": 12,
            "gp2Scale_sharded": True}
    gp = GP(x, y, init_hyperparameters=th, noise_variances=noise, gp2Scale=True, linalg_mode="sparseCGpre", args=args)
    th2 = th * np.array([1.0, 1.05, 0.97, 1.02])
    lml = gp.log_likelihood(th2)
    info = dict(gp.kv.last_sharded_sparse_info)
    ev = gp.kv.evaluate(th2, gp.likelihood.V, gp.prior.m)
    K = ev.csr.to_scipy()
    alpha = ev.KVinvY[:, 0].copy()
    iters = list(ev.info["cg_iters"])
    # the same evaluation on this rank alone (no sharding)
    args1 = dict(args, gp2Scale_sharded=False)
    gp1 = GP(x, y, init_hyperparameters=th, noise_variances=noise, gp2Scale=True, linalg_mode="sparseCGpre", args=args1)
    lml1 = gp1.log_likelihood(th2)
    ev1 = gp1.kv.evaluate(th2, gp1.likelihood.V, gp1.prior.m)
    K1 = ev1.csr.to_scipy()
    q.put((rank, lml, lml1, bool(np.array_equal(K.indptr, K1.indptr) and np.array_equal(K.indices, K1.indices)),
           bool(np.array_equal(K.data, K1.data)), float(np.max(np.abs(alpha - ev1.KVinvY[:, 0])) / np.max(np.abs(alpha))),
           iters, list(ev1.info["cg_iters"]), info))
    dist.barrier()
    dist.destroy_process_group()
    del ops, L


def test_sharded_sparse_two_ranks_nccl():
    """gp2Scale with CSR row slabs, row-sharded PCG (fvgp_pcg_sharded over NCCL) and split SLQ probes on 2 GPUs: the
    assembled matrix is bit-identical to the single-GPU one, the solution agrees to the CG tolerance and the LML --
    same probe stream -- to 1e-9."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + os.getpid() % 200
    procs = [ctx.Process(target=_sparse_worker, args=(r, 2, port, 40000, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, lml, lml1, same_pattern, same_values, dalpha, iters, iters1, info in res:
        assert same_pattern and same_values, info
        assert dalpha <= 1e-7 and abs(iters[0] - iters1[0]) <= 0.01 * iters1[0] + 2, (dalpha, iters, iters1)
        assert abs(lml / lml1 - 1) <= 1e-9, (lml, lml1)
        assert sum(info["nnz_per_rank"]) == info["nnz"] and info["rows"][1] % 32 == 0
    assert res[0][1] == res[1][1]                                         # both ranks return the same LML
