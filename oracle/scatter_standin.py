"""Restatement of the torch_scatter ops the reference calls (oracle; test infrastructure).

torch_scatter (PyG) is an un-vendored dependency of the reference, unpinned (install.sh:125
installs the newest wheel for torch 1.7.1 => 2.0.5-2.0.7) and not installable here, so its
published semantics are restated:
  * segment_csr(src, indptr, out=None, reduce): reduction along dim 0 over consecutive rows
    [indptr[i], indptr[i+1]); trailing dims broadcast; EMPTY segment -> 0 for every reduce
    (the reference relies on it: pooling.py:870); mean divides by max(count, 1); max/min
    return the first arg in segment order and route the gradient to that element only;
    sum/mean backward = gather (/count).
  * scatter_max / scatter_min(src, index, dim=0, dim_size) -> (values, arg) with
    arg == src.size(0) for empty groups (pooling.py:136 documents "-1 or n_points").
  * scatter(reduce='sum'|'mean'), scatter_add, scatter_mean.
The same functions are the oracle's own segment primitives (oracle/pooling_oracle.py).
Call sites in the reference: pooling.py:63,114,137,289,295,519,525,628,787,807,851;
image.py:1767,1867-1868,2240; modules.py:225; visibility.py:1264.
"""
from typing import Optional, Tuple

import torch


def _dense_index(indptr: torch.Tensor) -> torch.Tensor:
    n = indptr.numel() - 1
    counts = indptr[1:] - indptr[:-1]
    return torch.arange(n, device=indptr.device).repeat_interleave(counts)


def _first_arg(src2: torch.Tensor, dense: torch.Tensor, n_seg: int, is_max: bool):
    """values [n_seg,K], first arg [n_seg,K] (n_items when empty) of a sorted-index reduction."""
    n_items, K = src2.shape
    if n_items == 0:
        return (torch.zeros((n_seg, K), dtype=src2.dtype, device=src2.device),
                torch.full((n_seg, K), n_items, dtype=torch.long, device=src2.device))
    idx2 = dense.view(-1, 1).expand(-1, K)
    init = torch.full((n_seg, K), float("-inf") if is_max else float("inf"), dtype=src2.dtype,
                      device=src2.device)
    vals = init.scatter_reduce(0, idx2, src2, reduce="amax" if is_max else "amin", include_self=True)
    pos = torch.arange(n_items, device=src2.device).view(-1, 1).expand(-1, K)
    hit = src2 == vals.gather(0, idx2)
    cand = torch.where(hit, pos, torch.full_like(pos, n_items))
    arg = torch.full((n_seg, K), n_items, dtype=torch.long, device=src2.device)
    arg = arg.scatter_reduce(0, idx2, cand, reduce="amin", include_self=True)
    empty = arg == n_items
    vals = torch.where(empty, torch.zeros_like(vals), vals)
    return vals, arg


class _SegReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, dense, n_seg, reduce):
        shape = src.shape
        K = 1
        for d in shape[1:]:
            K *= int(d)
        src2 = src.reshape(shape[0], K)
        ctx.reduce, ctx.shape, ctx.n_seg = reduce, shape, n_seg
        counts = torch.bincount(dense, minlength=n_seg)
        if reduce in ("sum", "mean"):
            out = torch.zeros((n_seg, K), dtype=src.dtype, device=src.device)
            out.index_add_(0, dense, src2)  # sequential on CPU: segment order
            if reduce == "mean":
                out = out / counts.clamp(min=1).to(src.dtype).view(-1, 1)
            ctx.save_for_backward(dense, counts)
            arg = None
        else:
            out, arg = _first_arg(src2, dense, n_seg, reduce == "max")
            ctx.save_for_backward(dense, arg)
        # NB: return fresh (non-view) tensors: the reference modifies segment_csr outputs in place
        # (pooling.py:808 add_, Gating pooling.py:705-711), which autograd forbids on views
        # created inside a custom Function.
        tail = tuple(shape[1:])
        out = out if out.shape == (n_seg,) + tail else out.reshape((n_seg,) + tail).clone()
        if arg is None:
            return out, torch.empty(0, dtype=torch.long)
        return out, arg.reshape((n_seg,) + tail)

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        shape = ctx.shape
        n_items = shape[0]
        K = 1
        for d in shape[1:]:
            K *= int(d)
        g2 = grad_out.reshape(ctx.n_seg, K)
        if ctx.reduce in ("sum", "mean"):
            dense, counts = ctx.saved_tensors
            if ctx.reduce == "mean":
                g2 = g2 / counts.clamp(min=1).to(g2.dtype).view(-1, 1)
            gsrc = g2.index_select(0, dense)
        else:
            dense, arg = ctx.saved_tensors
            gsrc = torch.zeros((n_items + 1, K), dtype=g2.dtype, device=g2.device)
            gsrc.scatter_(0, arg, g2)  # empty segments write into the dummy row n_items
            gsrc = gsrc[:n_items]
        return gsrc.reshape(shape), None, None, None


def segment_csr(src: torch.Tensor, indptr: torch.Tensor, out: Optional[torch.Tensor] = None,
                reduce: str = "sum") -> torch.Tensor:
    assert out is None
    assert reduce in ("sum", "add", "mean", "max", "min")
    reduce = "sum" if reduce == "add" else reduce
    dense = _dense_index(indptr)
    return _SegReduce.apply(src, dense, indptr.numel() - 1, reduce)[0]


def segment_csr_arg(src, indptr, reduce="max"):
    dense = _dense_index(indptr)
    return _SegReduce.apply(src, dense, indptr.numel() - 1, reduce)


def _scatter_sorted_or_not(src, index, dim_size, reduce):
    """Generic scatter along dim 0 (index need not be sorted)."""
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() > 0 else 0
    shape = src.shape
    K = 1
    for d in shape[1:]:
        K *= int(d)
    src2 = src.reshape(shape[0], K)
    if reduce in ("sum", "mean"):
        out = torch.zeros((dim_size, K), dtype=src.dtype, device=src.device)
        out = out.index_add(0, index, src2)
        if reduce == "mean":
            counts = torch.bincount(index, minlength=dim_size).clamp(min=1)
            if torch.is_floating_point(out):
                out = out / counts.to(out.dtype).view(-1, 1)
            else:
                out = out // counts.view(-1, 1)
        return out.reshape((dim_size,) + tuple(shape[1:])), None
    vals, arg = _first_arg(src2, index, dim_size, reduce == "max")
    return (vals.reshape((dim_size,) + tuple(shape[1:])), arg.reshape((dim_size,) + tuple(shape[1:])))


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim: int = 0,
                out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    assert dim == 0 and out is None
    return _scatter_sorted_or_not(src, index, dim_size, "max")


def scatter_min(src: torch.Tensor, index: torch.Tensor, dim: int = 0,
                out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    assert dim == 0 and out is None
    return _scatter_sorted_or_not(src, index, dim_size, "min")


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and out is None
    return _scatter_sorted_or_not(src, index, dim_size, "sum")[0]


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and out is None
    return _scatter_sorted_or_not(src, index, dim_size, "mean")[0]


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and out is None
    if src.dtype == torch.bool:  # modules.py:225 sums a bool mask
        return _scatter_sorted_or_not(src.long(), index, dim_size, "sum")[0]
    reduce = "sum" if reduce == "add" else reduce
    res = _scatter_sorted_or_not(src, index, dim_size, reduce)
    return res[0]
