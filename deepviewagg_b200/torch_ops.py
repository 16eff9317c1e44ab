"""`torch.ops.dva.*`: the segment operators as registered PyTorch operators ("TorchScript" row).

The reference's three helpers `segment_softmax_csr`, `gather_csr`, `segment_gather_csr` are `@torch.jit.script`
functions that compile their callee `torch_scatter.segment_csr` (pooling.py:758-856).  Code that scripts its own
functions around them can call these registered operators instead: `torch.ops.dva.segment_csr(src, ptr, "max")`
is callable from TorchScript, dispatches to the same autograd Functions as `deepviewagg_b200.ops` (composite
registration: autograd, AMP and the C ABI underneath are those of `ops.py`), and fails loudly on CPU tensors like
every other entry point.  Importing this module registers the operators once per process.
"""
import torch

from . import ops

_LIB = torch.library.Library("dva", "DEF")
_LIB.define("segment_csr(Tensor src, Tensor indptr, str reduce) -> Tensor")
_LIB.define("gather_csr(Tensor src, Tensor csr_idx, int n_items) -> Tensor")
_LIB.define("segment_gather_csr(Tensor src, Tensor csr_idx, str reduce) -> Tensor")
_LIB.define("segment_softmax_csr(Tensor src, Tensor csr_idx, float eps, bool scaling) -> Tensor")
_LIB.define("view_attention(Tensor x, Tensor compat, Tensor csr_idx, int num_groups, Tensor? idx, Tensor? gate_weight, "
            "Tensor? gate_bias, bool group_scaling, float eps, bool idx_is_permutation) -> (Tensor, Tensor, Tensor?)")


def _segment_csr(src, indptr, reduce):
    return ops.segment_csr(src, indptr, reduce=reduce)


def _gather_csr(src, csr_idx, n_items):
    return ops.gather_csr(src, csr_idx, n_items=n_items if n_items >= 0 else None)


def _segment_gather_csr(src, csr_idx, reduce):
    return ops.segment_gather_csr(src, csr_idx, reduce=reduce)


def _segment_softmax_csr(src, csr_idx, eps, scaling):
    return ops.segment_softmax_csr(src, csr_idx, eps=eps, scaling=scaling)


def _view_attention(x, compat, csr_idx, num_groups, idx, gate_weight, gate_bias, group_scaling, eps, idx_is_permutation):
    return ops.view_attention(x, compat, csr_idx, num_groups, idx=idx, gate_weight=gate_weight, gate_bias=gate_bias,
                              group_scaling=group_scaling, eps=eps, idx_is_permutation=idx_is_permutation)


for _name, _fn in (("segment_csr", _segment_csr), ("gather_csr", _gather_csr),
                   ("segment_gather_csr", _segment_gather_csr), ("segment_softmax_csr", _segment_softmax_csr),
                   ("view_attention", _view_attention)):
    _LIB.impl(_name, _fn, "CompositeImplicitAutograd")

OPERATORS = ("segment_csr", "gather_csr", "segment_gather_csr", "segment_softmax_csr", "view_attention")
