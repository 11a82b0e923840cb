"""world_size-2 gloo test of the multi-GPU host logic (proposal sharding, result gather, max-over-ranks)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from fvgp_b200 import parallel
    r, _, w = parallel.init(backend="gloo")
    assert (r, w) == (rank, world)

    class FakeGP:                     # stands in for the GPU evaluation; host logic only
        def log_likelihood(self, th):
            return float(np.sum(th))

        def neg_log_likelihood_gradient(self, th):
            return 2.0 * th
    class FakePopulationGP(FakeGP):   # a GP that offers the population entry point: each rank's share is ONE call
        calls = []

        class _ML:
            def __init__(self, outer):
                self.outer = outer

            def evaluate_population(self, T, with_gradient=False, component=0):
                self.outer.calls.append(len(T))
                return T.sum(axis=1), (2.0 * T if with_gradient else None)

        def __init__(self):
            self.marginal_likelihood = self._ML(self)
    thetas = np.arange(15.0).reshape(5, 3)
    table = parallel.evaluate_proposals(FakeGP(), thetas)
    pgp = FakePopulationGP()
    table_pop = parallel.evaluate_proposals(pgp, thetas)
    assert np.array_equal(table_pop, table) and pgp.calls == [len(parallel.shard_proposals(5, rank, world))]
    assert np.array_equal(parallel.evaluate_proposals(pgp, thetas, with_gradient=False)[:, 0], table[:, 0])
    tmax = parallel.max_over_ranks(1.0 + rank)
    parallel.barrier()
    q.put((rank, parallel.shard_proposals(5, rank, world), table, tmax))
    dist.destroy_process_group()


def test_two_rank_proposal_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    thetas = np.arange(15.0).reshape(5, 3)
    expect = np.column_stack([thetas.sum(1), 2 * thetas])
    shards = {}
    for rank, shard, table, tmax in res:
        assert np.array_equal(table, expect)
        assert tmax == 2.0
        shards[rank] = shard
    assert sorted(shards[0] + shards[1]) == list(range(5)) and not set(shards[0]) & set(shards[1])
