"""Data-parallel plumbing for the view-aggregation path (one process per GPU, NCCL).

The path shards by scene / batch item: every operator is per-point over that point's own views and
MMBatch / ImageBatch are plain concatenations (reference core/multimodal/data.py:180-204,
image.py:1617-1672), so ranks own disjoint samples and never exchange activations.  The only
collective is one bucketed all-reduce of the pool-parameter gradients per step.
The reference itself is single-process: this module is new work, kept deliberately thin.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world_size):
    """Sample ids of `rank`: {i : i mod world_size == rank} (round-robin keeps scene sizes balanced)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world_size))


def flatten_grads(params):
    """One contiguous fp32 bucket holding every existing gradient, in parameter order."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return None, []
    bucket = torch.cat([g.detach().reshape(-1).float() for g in grads])
    return bucket, grads


def allreduce_gradients(params, average=True, group=None):
    """Sum (or average) the gradients of `params` over all ranks with ONE all_reduce.

    The bucket of the view-pool parameters is ~160 KB, i.e. latency-bound: one
    launch instead of one per tensor.  Returns the number of elements reduced."""
    params = list(params)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    bucket, grads = flatten_grads(params)
    if bucket is None:
        return 0
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    if average:
        bucket /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(bucket[off:off + n].view_as(g).to(g.dtype))
        off += n
    return int(bucket.numel())


def max_over_ranks(value, device=None, group=None):
    """max over ranks of a python float (step time): the job is as slow as its slowest rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value, device=None, group=None):
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())
