"""Replica-ensemble bookkeeping for multi-GPU runs.

A single rigid-body system does not shard (BASELINE.json: "replicas only"): N GPUs
run N independent systems, one process per GPU, with NO collective on the data path.  The only
communication is the timing protocol: a barrier on both sides of the timed region and a MAX over ranks
of the per-rank elapsed time, done here so that it can be exercised on CPU with the gloo backend."""
from __future__ import annotations

import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def replica_seed(base_seed: int, rank: int) -> int:
    """Distinct, reproducible workload seed per replica."""
    return int(base_seed) + int(rank)


def barrier(dist=None):
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """MAX all-reduce of a scalar (elapsed time); identity for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def ensemble_throughput(bodies_this_rank: int, steps: int, elapsed_s_this_rank: float, dist=None, device=None) -> float:
    """Whole-job body-steps/s of the replica ensemble: all bodies of all ranks x steps / slowest rank's time."""
    total_bodies = sum_over_ranks(bodies_this_rank, dist, device)
    slowest = max_over_ranks(elapsed_s_this_rank, dist, device)
    return total_bodies * steps / slowest
