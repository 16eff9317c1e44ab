"""Data-parallel plumbing for the view-aggregation path (one process per GPU, NCCL).

The path shards by scene / batch item: every operator is per-point over that point's own views and
MMBatch / ImageBatch are plain concatenations (reference core/multimodal/data.py:180-204,
image.py:1617-1672), so ranks own disjoint samples and never exchange activations.  The only
collective is one bucketed all-reduce of the pool-parameter gradients per step.
The reference itself is single-process: this module is new work, kept deliberately thin.
BatchNorm batch statistics and running buffers of the pool MLPs are PER RANK (each replica
normalises over its own batch, exactly what the single-process reference does with that batch); they
are not synchronised across ranks.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Pin this process (CPU affinity + preferred memory node) to the NUMA node its GPU hangs off, so
    that pinned host buffers allocated afterwards are local to the GPU's PCIe root.  With 8 ranks on a
    2-socket box, unbound ranks place half of their staging buffers on the far socket and every
    host<->device copy crosses the inter-socket link (round 1: e2e scaled 0.47 at N = 8).
    Best effort: returns a dict describing what was done; never raises."""
    info = {"node": None, "cpus": None, "mempolicy": None}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["pci"] = bdf
        if node < 0:
            return info
        info["node"] = node
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        info["cpus"] = len(cpus)
        # set_mempolicy(MPOL_PREFERRED = 1, nodemask): future pages come from `node` when it has room
        try:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
            info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
        except Exception as e:  # pragma: no cover
            info["mempolicy"] = f"unavailable ({type(e).__name__})"
    except Exception as e:
        info["error"] = f"{type(e).__name__}: {e}"
    return info


def shard_indices(n_items, rank, world_size):
    """Sample ids of `rank`: {i : i mod world_size == rank} (round-robin keeps scene sizes balanced)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world_size))


def flatten_grads(params):
    """One contiguous fp32 bucket over EVERY parameter that requires grad, in parameter order, zeros
    where a parameter has no gradient on this rank -- the layout must not depend on which gradients
    happen to exist locally (a rank whose batch has no modality data skips the pools,
    modules.py:148-160; unused parameters such as E_mix never get a .grad): ranks calling all_reduce
    with different element counts hang NCCL or average misaligned gradients."""
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return None, []
    bucket = torch.cat([(p.grad.detach().reshape(-1).float() if p.grad is not None
                         else torch.zeros(p.numel(), dtype=torch.float32, device=p.device)) for p in ps])
    return bucket, ps


def allreduce_gradients(params, average=True, group=None):
    """Sum (or average) the gradients of `params` over all ranks with ONE all_reduce.

    The bucket of the view-pool parameters is ~160 KB, i.e. latency-bound: one
    launch instead of one per tensor.  Returns the number of elements reduced."""
    params = list(params)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    bucket, ps = flatten_grads(params)
    if bucket is None:
        return 0
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    if average:
        bucket /= dist.get_world_size(group)
    off = 0
    for p in ps:
        n = p.numel()
        g = bucket[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()            # another rank had a gradient for it
        else:
            p.grad.copy_(g)
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
