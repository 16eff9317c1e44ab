"""Lexicographic sort / unique helpers with the reference's names
(torch_points3d/utils/multimodal.py:36-94).  One code path: a composite int64 key
(key = sum_i a_i * prod_{j>i} (max_j + 1), utils/multimodal.py:97-155) sorted on the device the
tensors live on (CUB radix sort through torch.sort on CUDA).  Unlike the reference's np.argsort
(introsort) the sort is STABLE, so results are deterministic; they agree with the reference up to
the order of equal keys.
"""
import numpy as np
import torch

# Key expected to be used for multimodal mappings (utils/multimodal.py:10)
MAPPING_KEY = 'mapping_index'


def tensor_idx(idx, device=None):
    """int / list / slice / ndarray / bool mask -> LongTensor (utils/multimodal.py:13-33)."""
    if idx is None:
        idx = torch.zeros(0, dtype=torch.long)
    elif isinstance(idx, int):
        idx = torch.tensor([idx], dtype=torch.long)
    elif isinstance(idx, (list, tuple, range)):
        idx = torch.tensor(list(idx), dtype=torch.long)
    elif isinstance(idx, slice):
        idx = torch.arange(idx.stop)[idx]
    elif isinstance(idx, np.ndarray):
        idx = torch.from_numpy(idx)
    if idx.dtype == torch.bool:
        idx = torch.where(idx)[0]
    assert idx.dtype == torch.int64, f"Expected LongTensor but got {idx.dtype} instead."
    return idx if device is None else idx.to(device)


def composite_key(*args):
    """-> (key int64 [n], bases list[int]) for 1-D integer tensors of equal length."""
    assert len(args) > 0, "At least one tensor must be provided."
    args = [torch.from_numpy(a) if isinstance(a, np.ndarray) else a for a in args]
    dev = args[0].device
    args = [a.to(dev).long() for a in args]
    assert all(a.dim() == 1 and a.shape == args[0].shape for a in args), \
        'All input tensors must be 1D and have the same shape.'
    if args[0].numel() == 0:
        return torch.zeros(0, dtype=torch.long, device=dev), [1] * len(args)
    maxs = torch.stack([a.abs().max() + 1 for a in args]).tolist()   # one host sync for all maxima
    total = 1
    for m in maxs:
        total *= int(m)
    assert total < torch.iinfo(torch.int64).max, 'composite key overflows int64'
    bases = []
    for i in range(len(args)):
        b = 1
        for m in maxs[i + 1:]:
            b *= int(m)
        bases.append(b)
    key = args[0] * bases[0]
    for a, b in zip(args[1:], bases[1:]):
        key = key + a * b
    return key, bases


def _restore(key, bases, dtypes):
    out = []
    for b, dt in zip(bases, dtypes):
        out.append((key // b).to(dt))
        key = key % b
    return out


def lexargsort(*args, use_cuda=False):
    """Indices sorting the inputs in lexicographic order (stable)."""
    key, _ = composite_key(*args)
    return torch.sort(key, stable=True).indices


def lexsort(*args, use_cuda=False):
    key, bases = composite_key(*args)
    out = _restore(torch.sort(key, stable=True).values, bases, [a.dtype for a in args])
    return out if len(out) > 1 else out[0]


def lexunique(*args, use_cuda=False):
    key, bases = composite_key(*args)
    out = _restore(torch.unique(key, sorted=True), bases, [a.dtype for a in args])
    return out if len(out) > 1 else out[0]


def lexargunique(*args, use_cuda=False):
    """Index of the FIRST occurrence of every unique key, in key order
    (np.unique(return_index=True) semantics of the reference's CPU path, utils/multimodal.py:307-311)."""
    key, _ = composite_key(*args)
    if key.numel() == 0:
        return key
    s = torch.sort(key, stable=True)
    first = torch.ones_like(s.values, dtype=torch.bool)
    first[1:] = s.values[1:] != s.values[:-1]
    return s.indices[first]
