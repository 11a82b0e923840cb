"""Multi-GPU plumbing: one process per GPU under torchrun, torch.distributed for the rendezvous.

What shards on this path: hyperparameter proposals are independent evaluations of
the same data (MCMC / differential-evolution / hgdl populations, gp_training.py:60-162), so N
GPUs evaluate N proposals concurrently with NO data-path collective -- "replicas"; the only
exchange is a gather of (H+1) doubles per proposal.  Paths whose DATA is sharded live elsewhere:
the 2-D block-cyclic dense factorisation in `sharded.py`, the row-sharded gp2Scale assembly /
probe-split log-determinant in `sharded_sparse.py`."""
import os

import numpy as np


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def shard_proposals(num, rank, world):
    """Indices of the proposals rank `rank` evaluates (round-robin)."""
    return list(range(rank, num, world))


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local, num, width):
    """All ranks receive the (num, width) table of per-proposal results; rows a rank did not
    evaluate are filled from their owner (sum of disjoint contributions)."""
    import torch
    import torch.distributed as dist
    table = np.zeros((num, width))
    for idx, row in local.items():
        table[idx] = row
    if dist.is_available() and dist.is_initialized():
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.as_tensor(table, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        table = t.cpu().numpy()
    return table


def evaluate_proposals(gp, thetas, with_gradient=True):
    """Evaluate LML (+ gradient) for every row of `thetas`, sharded over the ranks; each rank pushes its share through
    the population entry point (concurrent streams, one synchronisation) when the GP offers it."""
    rank, _, world = dist_env()
    thetas = np.asarray(thetas, dtype=np.float64)
    H = thetas.shape[1]
    mine = shard_proposals(len(thetas), rank, world)
    local = {}
    ml = getattr(gp, "marginal_likelihood", None)
    if mine and ml is not None and hasattr(ml, "evaluate_population"):
        lml, grad = ml.evaluate_population(thetas[mine], with_gradient=with_gradient)
        for k, i in enumerate(mine):
            row = np.zeros(1 + H)
            row[0] = lml[k]
            if with_gradient:
                row[1:] = grad[k]
            local[i] = row
        return gather_results(local, len(thetas), 1 + H)
    for i in mine:
        row = np.zeros(1 + H)
        row[0] = gp.log_likelihood(thetas[i])
        if with_gradient:
            row[1:] = gp.neg_log_likelihood_gradient(thetas[i])
        local[i] = row
    return gather_results(local, len(thetas), 1 + H)
