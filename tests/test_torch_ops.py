"""torch.ops.dva.*: registration, TorchScript-callability (the reference's helpers are scripted,
pooling.py:758-856) and equality with deepviewagg_b200.ops."""
import pytest
import torch

import deepviewagg_b200.torch_ops as torch_ops


@torch.jit.script
def _scripted_pool(x: torch.Tensor, ptr: torch.Tensor) -> torch.Tensor:
    m = torch.ops.dva.segment_csr(x, ptr, "max")
    a = torch.ops.dva.segment_softmax_csr(x, ptr, 1e-12, True)
    return torch.ops.dva.gather_csr(m, ptr, x.shape[0]) * a + torch.ops.dva.segment_gather_csr(x, ptr, "mean")


def test_operators_are_registered_and_scriptable():
    for name in torch_ops.OPERATORS:
        assert hasattr(torch.ops.dva, name), name
    schema = str(torch.ops.dva.segment_csr.default._schema)
    assert "Tensor src" in schema and "str reduce" in schema
    assert "dva::segment_csr" in str(_scripted_pool.graph)
    # CPU tensors fail loudly, like every other entry point (no fallback)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        torch.ops.dva.segment_csr(torch.zeros(3, 2), torch.tensor([0, 1, 3]), "sum")


@pytest.mark.gpu
def test_registered_operators_equal_ops_and_differentiate():
    from deepviewagg_b200 import ops
    gen = torch.Generator().manual_seed(5)
    counts = torch.randint(0, 6, (500,), generator=gen)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]).cuda()
    V = int(ptr[-1])
    x = torch.randn(V, 8, generator=gen).cuda()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = _scripted_pool(xa, ptr)
    m = ops.segment_csr(xb, ptr, reduce="max")
    yb = ops.gather_csr(m, ptr, n_items=V) * ops.segment_softmax_csr(xb, ptr, scaling=True) + \
        ops.segment_gather_csr(xb, ptr, reduce="mean")
    assert torch.equal(ya, yb)
    w = torch.randn(V, 8, generator=gen).cuda()
    ga, = torch.autograd.grad((ya * w).sum(), xa)
    gb, = torch.autograd.grad((yb * w).sum(), xb)
    assert torch.equal(ga, gb)
    # the fused pair through the registered operator
    G = 4
    compat = torch.randn(V, G, generator=gen).cuda()
    out_a, att_a, _ = torch.ops.dva.view_attention(x, compat, ptr, G, None, None, None, True, 1e-12, False)
    out_b, att_b, _ = ops.view_attention(x, compat, ptr, G, group_scaling=True)
    assert torch.equal(out_a, out_b) and torch.equal(att_a, att_b)
