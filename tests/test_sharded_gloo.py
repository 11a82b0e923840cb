"""Block-cyclic dense LML + gradient (fvgp_b200/sharded.py) on CPU with gloo: the choreography -- layout,
panel exchange, trailing updates, blocked solves, distributed TRTRI / LAUUM, sharded traces -- is executed for
several process grids with a torch-CPU LocalOps (tests/cpu_local_ops.py) and compared with the oracle."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(n, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3))
    y = np.sin(5 * x[:, 0]) * np.cos(3 * x[:, 1]) + x[:, 2] + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2) + 1e-3 * rng.random(n), np.array([1.1, .3, .45, .5])


def _worker(rank, world, port, grid, n, nb, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.set_num_threads(2)
    if world > 1:
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    from cpu_local_ops import CpuLocalOps
    from fvgp_b200 import sharded
    x, y, noise, theta = _problem(n)
    ev = sharded.ShardedDenseEvaluator(x, y, noise, nb=nb, grid=grid, ops=CpuLocalOps())
    out = ev.evaluate(0, theta[0], 1.0 / theta[1:], 1.0, np.full(n, y.mean()), want_gradient_theta=theta)
    q.put((rank, out["lml"], out["alpha"][:, 0], out["logdet"], out["traces"], ev.comm.bytes_received,
           ev.A.local_bytes()))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _run(grid, n, nb):
    world = grid[0] * grid[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() * 7 + grid[0] * 31 + grid[1] * 17 + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, n, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


def _reference(n):
    sys.path.insert(0, ROOT)
    from oracle import fvgp_oracle as orc
    x, y, noise, theta = _problem(n)
    lml = orc.dense_log_likelihood(x, y, theta, noise)
    grad = orc.dense_neg_log_likelihood_gradient(x, y, theta, noise, economical=True)
    return lml, grad


@pytest.mark.parametrize("grid,n,nb", [((1, 1), 700, 256), ((1, 2), 700, 128), ((2, 1), 650, 128),
                                       ((2, 2), 1000, 128), ((2, 2), 512, 128), ((1, 2), 100, 128),
                                       ((2, 4), 1100, 128)])
def test_block_cyclic_lml_and_gradient(grid, n, nb):
    res = _run(grid, n, nb)
    lml_ref, grad_ref = _reference(n)
    for rank, lml, alpha, logdet, traces, recv, local in res:
        assert abs(lml / lml_ref - 1) < 1e-10, (rank, lml, lml_ref)
        grad = 0.5 * traces                      # default mean: grad(-LML)_h = 1/2 sum (KV^-1 - b b^T) o dK_h
        assert np.max(np.abs(grad - grad_ref) / np.abs(grad_ref)) < 1e-8, (rank, grad, grad_ref)
    # every rank returns the same replicated results
    for r in res[1:]:
        assert r[1] == res[0][1] and np.array_equal(r[2], res[0][2])
    # storage: the lower staircase only, split over the grid
    total = sum(r[6] for r in res)
    nblk = -(-n // nb)
    assert total <= 8 * nb * nb * (nblk * (nblk + 1) // 2 + nblk)


def test_layout_arithmetic():
    sys.path.insert(0, ROOT)
    from fvgp_b200.sharded import BlockCyclicLayout, choose_grid, default_block
    assert [choose_grid(w) for w in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    n, nb = 1000, 128
    seen = np.zeros(n, dtype=int)
    for rank in range(8):
        lay = BlockCyclicLayout(n, nb, 2, 4, rank)
        if lay.q == 0:
            seen[lay.global_rows()] += 1
        assert lay.mloc() == len(lay.global_rows())
        for I in lay.row_blocks():
            assert lay.rows_from(I) == lay.lrow(I)
            assert lay.owner(I, lay.q) == rank
        assert lay.rows_from(lay.nblk) == lay.mloc()
    assert np.all(seen == 1)
    assert default_block(200_000, 8) == 2048 and default_block(4000, 2) % 128 == 0


def _worker_radial(rank, world, port, grid, n, nb, kind, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.set_num_threads(2)
    if world > 1:
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    from cpu_local_ops import CpuLocalOps
    from fvgp_b200 import sharded
    x, y, noise, _ = _problem(n)
    amp, inv, length = 1.2, np.array([2.5, 1.9, 2.2]), 0.8
    ev = sharded.ShardedDenseEvaluator(x, y, noise, nb=nb, grid=grid, ops=CpuLocalOps())
    out = ev.evaluate(kind, amp, inv, length, np.full(n, y.mean()))
    T = ev.gradient_traces(None, out["alpha_dev"][0], radial=(kind, amp, inv, length))
    q.put((rank, out["lml"], out["alpha"][:, 0], T))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("grid,kind", [((1, 2), 2), ((2, 2), 1), ((2, 1), 3), ((1, 1), 0)])
def test_block_cyclic_traces_of_the_radial_families(grid, kind):
    """Gradient traces against the fused descriptor (amp, inv_scale, length) of squared-exponential / Matern-5/2 /
    exponential / Matern-3/2 kernels on the sharded matrix vs central differences of the dense kernel matrix."""
    n, nb = 420, 128
    world = grid[0] * grid[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 25100 + (os.getpid() * 11 + grid[0] * 37 + grid[1] * 53 + kind) % 2000
    procs = [ctx.Process(target=_worker_radial, args=(r, world, port, grid, n, nb, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y, noise, _ = _problem(n)
    amp, inv, length = 1.2, np.array([2.5, 1.9, 2.2]), 0.8

    def K(amp, inv, length):
        d = np.sqrt((((x[:, None, :] - x[None, :, :]) * inv) ** 2).sum(-1))
        u = d / length
        if kind == 2:
            return amp * np.exp(-u ** 2 / 2)
        if kind == 0:
            return amp * (1 + np.sqrt(3) * u) * np.exp(-np.sqrt(3) * u)
        if kind == 1:
            return amp * (1 + np.sqrt(5) * u + 5 * u ** 2 / 3) * np.exp(-np.sqrt(5) * u)
        return amp * np.exp(-u)
    KV = K(amp, inv, length) + np.diag(noise)
    b = np.linalg.solve(KV, y - y.mean())
    M = np.linalg.inv(KV) - np.outer(b, b)
    eps = 1e-6
    fd = [np.sum(M * (K(amp + eps, inv, length) - K(amp - eps, inv, length))) / (2 * eps)]
    for i in range(3):
        e = np.zeros(3)
        e[i] = eps
        fd.append(np.sum(M * (K(amp, inv + e, length) - K(amp, inv - e, length))) / (2 * eps))
    fd.append(np.sum(M * (K(amp, inv, length + eps) - K(amp, inv, length - eps))) / (2 * eps))
    for rank, lml, alpha, T in res:
        assert np.max(np.abs(alpha - b)) <= 1e-8 * np.max(np.abs(b))
        assert np.max(np.abs(T - np.array(fd)) / np.maximum(np.abs(fd), 1.0)) <= 2e-6, (T, fd)
