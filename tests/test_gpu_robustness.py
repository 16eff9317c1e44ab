"""Boundary behaviour added in round 2: autocast, index
validation, pointers that do not cover every row, odd projection widths, duplicated row_index."""
import pytest
import torch

from conftest import load_golden
from oracle import pooling_oracle as O
from test_gpu_parity import _module_from_fixture, close, ragged_ptr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("amp_dtype,tol", [(torch.bfloat16, 3e-2), (torch.float16, 4e-3)])
def test_group_pool_under_autocast(amp_dtype, tol):
    """torch.autocast around the whole pool: the tensor-core projections are computed from fp32
    operands (custom_fwd cast_inputs), the feature operators run in the dtype they are handed; forward
    and every gradient stay within half-precision distance of the fp32 run, backward runs under the
    forward's autocast state (custom_bwd)."""
    from deepviewagg_b200.modules.multimodal.pooling import GroupBimodalCSRPool
    g = load_golden("group_pool_c64")
    m, _ = _module_from_fixture(g, GroupBimodalCSRPool)
    m.train()
    ptr, w = g["ptr"].cuda(), g["w"].cuda()

    def run(autocast, x_dtype):
        m.load_state_dict(g["sd"])
        x_mod = g["x_mod"].cuda().to(x_dtype).requires_grad_(True)
        x_map = g["x_map"].cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=amp_dtype, enabled=autocast):
            out = m(None, x_mod, x_map, ptr)
        grads = torch.autograd.grad((out.float() * w).sum(), [x_mod, x_map] + list(m.parameters()), allow_unused=True)
        return out, grads

    ref_out, ref_g = run(False, torch.float32)
    out, got_g = run(True, torch.float32)                 # fp32 activations under autocast
    assert torch.isfinite(out).all()
    close(out.float(), ref_out, tol, "autocast out")
    for a, b in zip(got_g, ref_g):
        if b is not None:
            assert a is not None and a.dtype == b.dtype
            close(a.float(), b.float(), 10 * tol, "autocast grad")
    out_h, got_h = run(True, amp_dtype)                   # half activations from an autocast CNN
    assert torch.isfinite(out_h).all() and got_h[0].dtype == amp_dtype
    close(out_h.float(), ref_out, 3 * tol, "autocast half-input out")


def test_gather_pool_index_validation_and_clamping():
    from deepviewagg_b200 import ops
    B, C, H, W = 2, 8, 6, 5
    fmap = torch.randn(B, H, W, C, device="cuda")
    aptr = torch.arange(4, device="cuda")
    images = torch.tensor([0, 1, 5], device="cuda")                      # image id 5 >= B
    pixels = torch.tensor([[1, 2], [4, 5], [9, 1]], dtype=torch.int32, device="cuda")  # x = 9 >= W
    out = ops.gather_pool(fmap.requires_grad_(True), images, pixels, aptr, "max", channels_last=True)
    # memory-safe: out-of-range indices are clamped into the map (forward read and backward write)
    assert torch.equal(out[0], fmap[0, 2, 1]) and torch.equal(out[2], fmap[1, 1, 4])
    out.sum().backward()
    assert torch.isfinite(fmap.grad).all() and float(fmap.grad.sum()) == pytest.approx(3 * C)
    ops.set_index_checks(True)
    try:
        with pytest.raises(IndexError):
            ops.gather_pool(fmap.detach(), images, pixels, aptr, "max", channels_last=True)
        ok = ops.gather_pool(fmap.detach(), images.clamp(max=1), pixels.clamp(max=4), aptr, "max", channels_last=True)
        assert ok.shape == (3, C)
    finally:
        ops.set_index_checks(False)


def test_uncovered_rows_are_zero_not_garbage():
    """csr_idx[0] > 0 or csr_idx[-1] < n_items (accepted by torch_scatter): rows outside every
    segment get 0 in gather_csr / segment_softmax_csr outputs and in the gradient of segment_csr."""
    from deepviewagg_b200 import ops
    n_items, K = 50, 12
    ptr = torch.tensor([5, 9, 9, 30], device="cuda")
    junk = torch.full((4096, K), float("nan"), device="cuda")           # poison the allocator's free blocks
    del junk
    src = torch.randn(n_items, K, device="cuda", requires_grad=True)
    for red in ("sum", "mean", "max", "min"):
        out = ops.segment_csr(src, ptr, reduce=red)
        (g,) = torch.autograd.grad(out.sum(), src)
        assert torch.isfinite(g).all() and (g[:5] == 0).all() and (g[30:] == 0).all() and (g[5:9] != 0).any()
    sm = ops.segment_softmax_csr(src.detach(), ptr)
    assert (sm[:5] == 0).all() and (sm[30:] == 0).all() and torch.allclose(sm[5:9].sum(0), torch.ones(K, device="cuda"))
    seg = torch.randn(3, K, device="cuda")
    gat = ops.gather_csr(seg, ptr, n_items=n_items)
    assert (gat[:5] == 0).all() and (gat[30:] == 0).all() and torch.equal(gat[9:30], seg[2].expand(21, K))


@pytest.mark.parametrize("K,N", [(130, 66), (65, 128), (96, 70), (33, 200)])
def test_linear_odd_widths_run_on_our_kernels(K, N):
    from deepviewagg_b200 import _lib, ops
    x = torch.randn(3000, K, device="cuda", requires_grad=True)
    w = torch.randn(N, K, device="cuda", requires_grad=True)
    n0 = _lib.launch_count()
    z = ops.linear(x, w)
    assert _lib.launch_count() > n0 and z.shape == (3000, N)          # no library GEMM fallback
    gz = torch.randn_like(z)
    gx, gw = torch.autograd.grad(z, [x, w], gz)
    ref = x.double() @ w.double().t()
    close(z.double(), ref, 3e-6, "padded linear")
    close(gx.double(), gz.double() @ w.double(), 3e-6, "padded dX")
    close(gw.double(), gz.double().t() @ x.double(), 3e-6, "padded dW")
    with pytest.raises(RuntimeError):
        ops.linear(x.cpu(), w.cpu())


def test_row_index_with_duplicates_accumulates():
    """a caller-supplied row_index that is not a permutation must take the accumulating
    backward (each x_mod row receives the sum over the views that read it)."""
    from deepviewagg_b200.modules.multimodal.pooling import GroupBimodalCSRPool
    g = load_golden("group_pool_c64")
    m, _ = _module_from_fixture(g, GroupBimodalCSRPool)
    m.eval()
    x_mod, x_map, ptr = g["x_mod"].cuda(), g["x_map"].cuda(), g["ptr"].cuda()
    V = x_map.shape[0]
    ridx = torch.randint(0, V // 3, (V,), generator=torch.Generator().manual_seed(3)).cuda()   # many duplicates
    xa = x_mod.clone().requires_grad_(True)
    out_a = m(None, xa, x_map, ptr, row_index=ridx)
    xb = x_mod.clone().requires_grad_(True)
    out_b = m(None, xb[ridx], x_map, ptr)
    close(out_a, out_b, 1e-6, "dup row_index out")
    w = torch.randn_like(out_a)
    (ga,) = torch.autograd.grad((out_a * w).sum(), xa)
    (gb,) = torch.autograd.grad((out_b * w).sum(), xb)
    close(ga, gb, 1e-5, "dup row_index grad")
    m2, _ = _module_from_fixture(g, GroupBimodalCSRPool)            # save_last taps the gathered rows
    m2.eval()
    m2(None, x_mod, x_map, ptr, row_index=ridx)
    assert torch.equal(m2._last_x_mod, m2.E_mod(x_mod)[ridx])
