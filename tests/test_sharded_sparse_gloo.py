"""Multi-GPU gp2Scale (fvgp_b200/sharded_sparse.py) on CPU with gloo: slab edges, count balance, in-place all-gather of
the row strips, row-sharded CG contract and probe split, with a numpy/torch-CPU ops stand-in (tests/cpu_sparse_ops.py)
against the oracle's single-process assembly (gp2Scale_covariance.py:313-431 restated in oracle/fvgp_oracle.py)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(n, clustered, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 2))
    if clustered:                                      # a dense cluster in the first rows: equal rows != equal work
        x[:n // 4] = 0.5 + 0.03 * rng.standard_normal((n // 4, 2))
    y = np.sin(5 * x[:, 0]) + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2) + 1e-3 * rng.random(n), np.array([1.3, .09, .11])


def _worker(rank, world, port, n, clustered, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.set_num_threads(2)
    if world > 1:
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    from cpu_sparse_ops import CpuSparseOps
    from fvgp_b200 import sharded_sparse
    x, y, noise, theta = _problem(n, clustered)
    ops = CpuSparseOps()
    ev = sharded_sparse.ShardedSparseEvaluator(ops=ops)
    csr, rows = ev.assemble(torch.as_tensor(x), theta, torch.as_tensor(noise))
    K = csr.to_scipy()
    b = torch.as_tensor(y - y.mean())
    sol, info, iters, relres = ev.pcg(csr, b, rtol=1e-10)
    est, var, samples = ev.slq_logdet(csr, 12, 7, seed=3)
    q.put((rank, K.indptr.copy(), K.indices.copy(), K.data.copy(), list(rows), dict(ev.info), sol.numpy().copy(), info,
           iters, est, samples.copy(), ops.count_calls))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _run(world, n, clustered):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 27100 + (os.getpid() * 13 + world * 101 + n + int(clustered)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, clustered, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


@pytest.mark.parametrize("world,n,clustered", [(1, 300, False), (2, 500, False), (3, 450, True), (4, 333, True)])
def test_row_sharded_assembly_solve_and_logdet(world, n, clustered):
    sys.path.insert(0, ROOT)
    import scipy.sparse.linalg as spla
    from oracle import fvgp_oracle as orc
    res = _run(world, n, clustered)
    x, y, noise, theta = _problem(n, clustered)
    Kref = orc.add_kv(orc.gp2scale_covariance(x, x, theta, batch=97, symmetric=True), noise)
    Kref.sort_indices()
    sol_ref = spla.spsolve(Kref.tocsc(), y - y.mean())
    single = _run(1, n, clustered)[0] if world > 1 else res[0]
    for rank, indptr, indices, data, rows, info, sol, cg_info, iters, est, samples, count_calls in res:
        assert np.array_equal(indptr, Kref.indptr) and np.array_equal(indices, Kref.indices)     # bit-exact pattern
        assert np.array_equal(data, Kref.data)                                                   # same kernel, same values
        assert rows[0] == 0 and rows[-1] == n and all(r % 32 == 0 for r in rows[1:-1]) and rows == sorted(rows)
        assert sum(info["nnz_per_rank"]) == Kref.nnz == info["nnz"]
        assert cg_info == 0 and np.max(np.abs(sol - sol_ref)) <= 1e-8 * np.max(np.abs(sol_ref))
        assert np.array_equal(samples, single[10]) and est == single[9]         # probe split = the single-rank stream
        assert count_calls == (2 if info["rebalanced"] else 1)
    if clustered and world > 1:
        info = res[0][5]
        assert info["rebalanced"]
        assert max(info["nnz_per_rank"]) <= 1.35 * info["nnz"] / world          # edges are multiples of 32 rows
    for r in res[1:]:
        assert r[4] == res[0][4] and np.array_equal(r[6], res[0][6])            # every rank: same slabs, same solution


def test_slab_arithmetic():
    sys.path.insert(0, ROOT)
    from fvgp_b200.sharded_sparse import balanced_slabs, equal_slabs, split_probes
    assert equal_slabs(1000, 4) == [0, 256, 512, 768, 1000]
    assert equal_slabs(50, 4) == [0, 32, 50, 50, 50]
    assert equal_slabs(1_000_000, 8)[1] % 32 == 0 and equal_slabs(1_000_000, 8)[-1] == 1_000_000
    counts = np.r_[np.full(256, 30), np.full(768, 10)]
    offs = balanced_slabs(np.cumsum(counts), 1024, 2)
    assert offs == [0, 256, 1024]                      # 7680 of 15360 entries sit in the first 256 rows
    offs = balanced_slabs(np.cumsum(np.ones(1000, dtype=int)), 1000, 3)
    assert offs[0] == 0 and offs[-1] == 1000 and all(o % 32 == 0 for o in offs[1:-1]) and offs == sorted(offs)
    assert split_probes(10, 4) == [0, 2, 5, 7, 10] and split_probes(3, 8)[-1] == 3
