"""Data-parallel plumbing of the INR path: pixels of a global batch are independent
given the parameters and every loss term is a batch mean, so each rank processes its own contiguous
shard and ONE all-reduce over the flat gradient buffer restores the single-process gradient.
Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
from typing import Tuple

import torch


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`; shards differ by at most one item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def allreduce_gradient(flat_grad: torch.Tensor, dist, world: int, local_count: int = 1, global_count: int = 0) -> float:
    """Sums `flat_grad` over ranks in place and returns the factor that turns the sum of per-rank
    batch-mean gradients into the global batch-mean gradient (applied by the optimiser's unscale, so no
    extra pass over the buffer): local_count / global_count, i.e. 1 / world for equal shards."""
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    if global_count:
        return float(local_count) / float(global_count)
    return 1.0 / float(world)
