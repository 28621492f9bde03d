"""Multi-GPU plumbing for the two ways the path shards:
 (1) the sample axis of a loss: contiguous index blocks per rank, per-rank partial sums already divided
     by the FULL N, combined with ONE all-reduce(sum) of a [K x (1+P)] double buffer per batch;
 (2) independent fits (material x model x metric): static cost-weighted partition, no collective.
Host-side only; the per-rank work is bbmcu_loss_create(first, count) + bbmcu_loss_eval."""
import numpy as np


def shard_range(n, rank, world):
    """contiguous block of [0, n) owned by `rank` (grid-index order keeps theta_h-major locality)"""
    per = (n + world - 1) // world
    first = min(n, rank * per)
    return first, max(0, min(per, n - first))


def interleaved_indices(n, rank, world, block=1024):
    """linearizer indices of the shard BBMCU_LOSS_SHARD_INTERLEAVED gives `rank`: blocks rank, rank + world, ... of `block`
    consecutive samples, in the shard's own order (bbmcu_loss_terms order).  Equal work per shard where contiguous ranges are not."""
    blocks = (n + block - 1) // block
    mine = np.arange(rank, blocks, world, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n]


def partition_by_cost(costs, world):
    """longest-processing-time-first assignment of independent jobs to ranks; returns a list of index lists"""
    order = np.argsort(-np.asarray(costs, np.float64), kind="stable")
    load = np.zeros(world)
    out = [[] for _ in range(world)]
    for j in order:
        r = int(np.argmin(load))
        out[r].append(int(j))
        load[r] += costs[j]
    return out


def all_reduce_sum(tensor, group=None):
    """sum a [K x (1+P)] float64 tensor over ranks (NCCL on GPU, gloo in the CPU tests)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor
