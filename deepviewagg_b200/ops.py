"""Autograd operators over the C ABI (include/dva_b200.h).

Each function mirrors one operator of the reference's multimodal path (same argument meaning,
same empty-segment / tie / eps semantics) and is backed ONLY by the sm_100a kernels of
libdva_b200.so: CPU tensors or a missing library raise.  Reference citations are relative to the
reference repository root.
"""
import math
import os

import torch

# torch.amp integration (SURVEY 8b "Autograd / AMP / recompute"): every autograd.Function's forward is
# wrapped in torch.amp.custom_fwd and its backward in custom_bwd, so that (a) the backward runs under
# the autocast state of its forward and (b) the tensor-core projection is computed from fp32 operands
# (cast_inputs) whatever dtype autocast hands it -- never less precise than the reference's fp16
# autocast path (models/segmentation/sparseconv3d.py:24).  The feature operators (segment / gather /
# attention) run in the dtype of their inputs (fp32, bf16 or fp16 storage, fp32 accumulation), which
# is what torch_scatter does under autocast.
_fwd = torch.amp.custom_fwd(device_type="cuda")
_fwd_f32 = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = torch.amp.custom_bwd(device_type="cuda")

from . import _lib
from ._lib import DTYPE_CODES, REDUCE_CODES, check, dtype_code, ptr, require_cuda, stream_ptr


def _as_2d(src):
    if src.dim() == 1:
        return src.contiguous().view(-1, 1)
    if src.dim() == 2:
        return src.contiguous()
    return src.contiguous().view(src.shape[0], -1)


def _check_csr(csr_idx, device):
    if csr_idx.dtype != torch.int64:
        raise TypeError("csr_idx must be a LongTensor (core/multimodal/csr.py:54)")
    if csr_idx.dim() != 1 or csr_idx.numel() < 1:
        raise ValueError("csr_idx must be a 1D pointer tensor of size n_groups + 1")
    if csr_idx.device != device:
        raise RuntimeError("csr_idx must live on the device of the features")
    return csr_idx.contiguous()


# --------------------------------------------------------------------------------------------
# segment_csr  (torch_scatter.segment_csr as used at pooling.py:63,289,295,519,525,628,787,807)
# --------------------------------------------------------------------------------------------
class _SegmentCSR(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, src, csr_idx, reduce):
        require_cuda(src, csr_idx)
        lib = _lib.load()
        code = REDUCE_CODES[reduce]
        shape = src.shape
        s2 = _as_2d(src)
        csr_idx = _check_csr(csr_idx, src.device)
        n_seg, n_items, K = csr_idx.numel() - 1, s2.shape[0], s2.shape[1]
        # outputs are allocated in their final shape: returning a view from a custom Function
        # would forbid the in-place updates the reference applies downstream (Gating,
        # pooling.py:705-711)
        out = torch.empty((n_seg,) + tuple(shape[1:]), dtype=src.dtype, device=src.device)
        arg = None
        if code in (2, 3):
            arg = torch.empty((n_seg, K), dtype=torch.int64, device=src.device)
        with torch.cuda.device(src.device):
            check(lib.dva_segment_csr_fwd(ptr(s2), ptr(csr_idx), ptr(out), ptr(arg), n_seg, n_items,
                                          K, code, dtype_code(s2), stream_ptr()), "dva_segment_csr_fwd")
        ctx.code, ctx.n_items, ctx.in_shape = code, n_items, shape
        ctx.save_for_backward(csr_idx, arg)
        return out

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        csr_idx, arg = ctx.saved_tensors
        lib = _lib.load()
        g2 = _as_2d(grad_out)
        n_seg, K = g2.shape
        gsrc = torch.empty(ctx.in_shape, dtype=g2.dtype, device=g2.device)
        with torch.cuda.device(g2.device):
            check(lib.dva_segment_csr_bwd(ptr(g2), ptr(csr_idx), ptr(arg), ptr(gsrc), n_seg,
                                          ctx.n_items, K, ctx.code, dtype_code(g2), stream_ptr()),
                  "dva_segment_csr_bwd")
        return gsrc, None, None


def segment_csr(src, indptr, out=None, reduce="sum"):
    """torch_scatter.segment_csr(src, indptr, out=None, reduce) along dim 0.

    Empty segments reduce to 0 for every mode (pooling.py:870); max/min route the gradient to
    the first arg-max/min row of the segment.
    """
    if out is not None:
        raise NotImplementedError("segment_csr(out=...) is not used by the reference path")
    if reduce not in REDUCE_CODES:
        raise ValueError(f"unknown reduce '{reduce}'")
    return _SegmentCSR.apply(src, indptr, reduce)


def segment_csr_arg(src, indptr, reduce="max"):
    """(values, first-arg rows) like torch_scatter.segment_max_csr; arg = n_items when empty."""
    require_cuda(src, indptr)
    lib = _lib.load()
    s2 = _as_2d(src)
    indptr = _check_csr(indptr, src.device)
    n_seg, n_items, K = indptr.numel() - 1, s2.shape[0], s2.shape[1]
    out = torch.empty((n_seg, K), dtype=src.dtype, device=src.device)
    arg = torch.empty((n_seg, K), dtype=torch.int64, device=src.device)
    with torch.cuda.device(src.device):
        check(lib.dva_segment_csr_fwd(ptr(s2), ptr(indptr), ptr(out), ptr(arg), n_seg, n_items, K,
                                      REDUCE_CODES[reduce], dtype_code(s2), stream_ptr()),
              "dva_segment_csr_fwd")
    tail = tuple(src.shape[1:])
    return out.view((n_seg,) + tail), arg.view((n_seg,) + tail)


# --------------------------------------------------------------------------------------------
# gather_csr (pooling.py:813-841)
# --------------------------------------------------------------------------------------------
class _GatherCSR(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, src, csr_idx, n_items):
        require_cuda(src, csr_idx)
        lib = _lib.load()
        s2 = _as_2d(src)
        csr_idx = _check_csr(csr_idx, src.device)
        n_seg, K = csr_idx.numel() - 1, s2.shape[1]
        out = torch.empty((n_items,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        with torch.cuda.device(src.device):
            check(lib.dva_gather_csr(ptr(s2), ptr(csr_idx), ptr(out), n_seg, n_items, K,
                                     dtype_code(s2), stream_ptr()), "dva_gather_csr")
        ctx.in_shape = src.shape
        ctx.save_for_backward(csr_idx)
        return out

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (csr_idx,) = ctx.saved_tensors
        lib = _lib.load()
        g2 = _as_2d(grad_out)
        n_items, K = g2.shape
        n_seg = csr_idx.numel() - 1
        gsrc = torch.empty(ctx.in_shape, dtype=g2.dtype, device=g2.device)
        with torch.cuda.device(g2.device):
            check(lib.dva_segment_csr_fwd(ptr(g2), ptr(csr_idx), ptr(gsrc), None, n_seg, n_items, K,
                                          REDUCE_CODES["sum"], dtype_code(g2), stream_ptr()),
                  "dva_segment_csr_fwd")
        return gsrc, None, None


def gather_csr(src, csr_idx, n_items=None):
    """Redistribute segment-level rows to their items (pooling.py:813-841).

    `n_items` avoids the device->host read of csr_idx[-1] when the caller knows V already.
    """
    if not torch.is_floating_point(src):
        raise ValueError("`gather_csr` can only be computed over tensors with floating point data types.")
    if csr_idx.dim() != 1:
        raise ValueError("`gather_csr` can only be computed over 1D CSR indices.")
    if src.dim() > 2:
        raise NotImplementedError("`gather_csr` can only be computed over 1D or 2D source tensors.")
    if n_items is None:
        n_items = int(csr_idx[-1].item())
    return _GatherCSR.apply(src, csr_idx, n_items)


def segment_gather_csr(src, csr_idx, reduce="sum"):
    """segment_csr then gather_csr (pooling.py:844-856)."""
    return gather_csr(segment_csr(src, csr_idx, reduce=reduce), csr_idx, n_items=src.shape[0])


# --------------------------------------------------------------------------------------------
# segment_softmax_csr (pooling.py:758-810)
# --------------------------------------------------------------------------------------------
class _SegmentSoftmaxCSR(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, src, csr_idx, eps, scaling):
        require_cuda(src, csr_idx)
        lib = _lib.load()
        s2 = _as_2d(src)
        csr_idx = _check_csr(csr_idx, src.device)
        n_seg, n_items, K = csr_idx.numel() - 1, s2.shape[0], s2.shape[1]
        out = torch.empty(src.shape, dtype=src.dtype, device=src.device)
        with torch.cuda.device(src.device):
            check(lib.dva_segment_softmax_csr_fwd(ptr(s2), ptr(csr_idx), ptr(out), n_seg, n_items, K,
                                                  float(eps), int(bool(scaling)), dtype_code(s2),
                                                  stream_ptr()), "dva_segment_softmax_csr_fwd")
        ctx.scaling, ctx.in_shape = bool(scaling), src.shape
        ctx.save_for_backward(csr_idx, out)
        return out

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        csr_idx, out = ctx.saved_tensors
        lib = _lib.load()
        g2 = _as_2d(grad_out)
        n_items, K = g2.shape
        gsrc = torch.empty(ctx.in_shape, dtype=g2.dtype, device=g2.device)
        with torch.cuda.device(g2.device):
            check(lib.dva_segment_softmax_csr_bwd(ptr(out), ptr(g2), ptr(csr_idx), ptr(gsrc),
                                                  csr_idx.numel() - 1, n_items, K, int(ctx.scaling),
                                                  dtype_code(g2), stream_ptr()),
                  "dva_segment_softmax_csr_bwd")
        return gsrc, None, None, None


def segment_softmax_csr(src, csr_idx, eps=1e-12, scaling=False):
    """Equivalent of scatter_softmax for CSR indices (pooling.py:758-810), same signature."""
    if not torch.is_floating_point(src):
        raise ValueError("`segment_csr_softmax` can only be computed over tensors with floating point data types.")
    if csr_idx.dim() != 1:
        raise ValueError("`segment_csr_softmax` can only be computed over 1D CSR indices.")
    if src.dim() > 2:
        raise NotImplementedError("`segment_csr_softmax` can only be computed over 1D or 2D source tensors.")
    return _SegmentSoftmaxCSR.apply(src, csr_idx, eps, scaling)


# --------------------------------------------------------------------------------------------
# fused view attention (modules.py:518 + pooling.py:285-300 / 515-530)
# --------------------------------------------------------------------------------------------
def _scatter_add_rows(src, idx, n_rows):
    """fp32 [n_rows, C] with dst[idx[v]] += src[v] (dva_scatter_add_rows)."""
    lib = _lib.load()
    src = src.contiguous()
    V, C = src.shape
    dst = torch.zeros((n_rows, C), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        check(lib.dva_scatter_add_rows(ptr(src), ptr(idx.contiguous()), ptr(dst), V, n_rows, C, dtype_code(src),
                                       stream_ptr()), "dva_scatter_add_rows")
    return dst


def fused_groups_supported(num_groups):
    return 1 <= num_groups <= 32 and (num_groups & (num_groups - 1)) == 0


class _ViewAttention(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, x, idx, compat, csr_idx, gate_w, gate_b, num_groups, group_scaling, eps,
                idx_is_permutation):
        require_cuda(x, idx, compat, csr_idx, gate_w, gate_b)
        lib = _lib.load()
        x = x.contiguous()
        compat = compat.float().contiguous()
        csr_idx = _check_csr(csr_idx, x.device)
        N, V, G = csr_idx.numel() - 1, compat.shape[0], int(num_groups)
        R, C = x.shape
        if compat.shape[1] != G:
            raise ValueError(f"compatibilities must be [V,{G}], got {tuple(compat.shape)}")
        idx64 = 0
        if idx is not None:
            if idx.dtype not in (torch.int32, torch.int64):
                raise TypeError("idx must be int32 or int64")
            idx = idx.contiguous()
            idx64 = int(idx.dtype == torch.int64)
            if idx.numel() != V:
                raise ValueError("idx must hold one row id per view")
        elif R != V:
            raise ValueError("x must hold one row per view when idx is None")
        gw = gate_w.detach().float().contiguous().view(-1) if gate_w is not None else None
        gb = gate_b.detach().float().contiguous().view(-1) if gate_b is not None else None
        out = torch.empty((N, C), dtype=x.dtype, device=x.device)
        att = torch.empty((V, G), dtype=torch.float32, device=x.device)
        seg_max = torch.empty((N, G), dtype=torch.float32, device=x.device)
        seg_den = torch.empty((N, G), dtype=torch.float32, device=x.device)
        seg_arg = torch.empty((N, G), dtype=torch.int32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.dva_view_attention_fwd(ptr(x), ptr(idx), idx64, ptr(compat), ptr(csr_idx), ptr(gw),
                                             ptr(gb), ptr(out), ptr(att), ptr(seg_max), ptr(seg_den),
                                             ptr(seg_arg), N, V, R, C, G, int(bool(group_scaling)),
                                             float(eps), dtype_code(x), stream_ptr()),
                  "dva_view_attention_fwd")
        ctx.cfg = (N, V, R, C, G, bool(group_scaling), idx64, bool(idx_is_permutation),
                   gate_w.shape if gate_w is not None else None,
                   gate_b.shape if gate_b is not None else None,
                   gate_w.dtype if gate_w is not None else None)
        ctx.save_for_backward(x, idx, compat, csr_idx, gw, gb, seg_max, seg_den, seg_arg)
        ctx.mark_non_differentiable(att, seg_max)
        return out, att, seg_max

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out, _ga, _gm):
        x, idx, compat, csr_idx, gw, gb, seg_max, seg_den, seg_arg = ctx.saved_tensors
        N, V, R, C, G, scaling, idx64, is_perm, w_shape, b_shape, w_dtype = ctx.cfg
        lib = _lib.load()
        grad_out = grad_out.contiguous()
        scatter = int(idx is not None and is_perm and R == V)
        gx_rows = torch.empty((V, C), dtype=x.dtype, device=x.device)
        gcompat = torch.empty((V, G), dtype=torch.float32, device=x.device)
        ggate, ws, ws_bytes = None, None, 0
        if gw is not None:
            ggate = torch.empty((2, G), dtype=torch.float32, device=x.device)
            ws_bytes = int(lib.dva_view_attention_bwd_workspace_bytes(G))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.dva_view_attention_bwd(ptr(x), ptr(idx), idx64, ptr(compat), ptr(csr_idx), ptr(gw),
                                             ptr(gb), ptr(grad_out), ptr(seg_max), ptr(seg_den),
                                             ptr(seg_arg), ptr(gx_rows), ptr(gcompat), ptr(ggate),
                                             scatter, N, V, R, C, G, int(scaling), dtype_code(x), ptr(ws),
                                             ws_bytes, stream_ptr()), "dva_view_attention_bwd")
        if idx is None or scatter:
            gx = gx_rows
        else:  # general (non-injective) gather: accumulate duplicated rows (red.global.add.v4.f32 kernel)
            gx = _scatter_add_rows(gx_rows, idx.long(), R).to(x.dtype)
        g_w = ggate[0].view(w_shape).to(w_dtype) if gw is not None else None
        g_b = ggate[1].view(b_shape).to(w_dtype) if gw is not None else None
        return gx, None, gcompat, None, g_w, g_b, None, None, None, None


def view_attention(x, compat, csr_idx, num_groups, idx=None, gate_weight=None, gate_bias=None,
                   group_scaling=False, eps=1e-12, idx_is_permutation=False):
    """Fused gather + group softmax + weighted sum (+ gating).

    Returns (x_pool [N,C], attentions [V,G], seg_max [N,G]) where
      attentions = segment_softmax_csr(compat, csr_idx, scaling=group_scaling)
      x_pool     = segment_csr(x[idx] * expand_group_feat(attentions), csr_idx, 'sum')
                   * expand_group_feat(tanh(relu(w * segment_csr(compat,'max') + b)))   if gating
    i.e. the chain modules.py:518 -> pooling.py:285-300. `idx` (int32/int64 [V], optional) is the
    row of `x` feeding each view (e.g. ImageData.view_cat_sorting, image.py:1549-1574).
    """
    if not fused_groups_supported(num_groups):
        raise NotImplementedError("fused view attention needs num_groups to be a power of two <= 32")
    if (gate_weight is None) != (gate_bias is None):
        raise ValueError("gate_weight and gate_bias go together")
    return _ViewAttention.apply(x, idx, compat, csr_idx, gate_weight, gate_bias, num_groups,
                                group_scaling, eps, idx_is_permutation)


# --------------------------------------------------------------------------------------------
# ragged Q.K compatibilities (pooling.py:499-512)
# --------------------------------------------------------------------------------------------
class _QKScores(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, keys, queries, csr_idx, num_groups, scale):
        require_cuda(keys, queries, csr_idx)
        lib = _lib.load()
        k32, q32 = keys.float().contiguous(), queries.float().contiguous()
        csr_idx = _check_csr(csr_idx, keys.device)
        N, V, G = csr_idx.numel() - 1, k32.shape[0], int(num_groups)
        D = k32.shape[1] // G
        if k32.shape[1] != G * D or q32.shape != (N, G * D):
            raise ValueError("keys must be [V,G*D] and queries [N,G*D]")
        compat = torch.empty((V, G), dtype=torch.float32, device=keys.device)
        with torch.cuda.device(keys.device):
            check(lib.dva_qk_scores_fwd(ptr(k32), ptr(q32), ptr(csr_idx), ptr(compat), N, V, G, D,
                                        float(scale), stream_ptr()), "dva_qk_scores_fwd")
        ctx.cfg = (N, V, G, D, float(scale), keys.dtype, queries.dtype)
        ctx.save_for_backward(k32, q32, csr_idx)
        return compat

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcompat):
        k32, q32, csr_idx = ctx.saved_tensors
        N, V, G, D, scale, kd, qd = ctx.cfg
        lib = _lib.load()
        gcompat = gcompat.float().contiguous()
        gk, gq = torch.empty_like(k32), torch.empty_like(q32)
        with torch.cuda.device(k32.device):
            check(lib.dva_qk_scores_bwd(ptr(k32), ptr(q32), ptr(csr_idx), ptr(gcompat), ptr(gk), ptr(gq),
                                        N, V, G, D, scale, stream_ptr()), "dva_qk_scores_bwd")
        return gk.to(kd), gq.to(qd), None, None, None


def qk_scores(keys, queries, csr_idx, num_groups, dim_scaling=True):
    """compat[v,g] = sum_d K[v,g,d] Q[point(v),g,d] (/ sqrt(D) if dim_scaling), pooling.py:499-512."""
    D = keys.shape[1] // num_groups
    scale = 1.0 / math.sqrt(D) if dim_scaling else 1.0
    return _QKScores.apply(keys, queries, csr_idx, num_groups, scale)


# --------------------------------------------------------------------------------------------
# heuristic pool (pooling.py:129-152)
# --------------------------------------------------------------------------------------------
class _HeuristicPool(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, x_mod, x_map, csr_idx, feat, use_max):
        require_cuda(x_mod, x_map, csr_idx)
        lib = _lib.load()
        x_mod = x_mod.contiguous()
        m32 = x_map.float().contiguous()
        csr_idx = _check_csr(csr_idx, x_mod.device)
        N, V, C = csr_idx.numel() - 1, x_mod.shape[0], x_mod.shape[1]
        out = torch.empty((N, C), dtype=x_mod.dtype, device=x_mod.device)
        arg = torch.empty((N,), dtype=torch.int64, device=x_mod.device)
        with torch.cuda.device(x_mod.device):
            check(lib.dva_heuristic_pool_fwd(ptr(x_mod), ptr(m32), m32.shape[1], int(feat), ptr(csr_idx),
                                             ptr(out), ptr(arg), N, V, C, int(bool(use_max)),
                                             dtype_code(x_mod), stream_ptr()), "dva_heuristic_pool_fwd")
        ctx.V = V
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (arg,) = ctx.saved_tensors
        # each point picks a distinct view (arg == V: none, skipped by the kernel)
        g = _scatter_add_rows(grad_out.contiguous(), arg, ctx.V).to(grad_out.dtype)
        return g, None, None, None, None


def heuristic_pool(x_mod, x_map, csr_idx, feat, mode="max"):
    return _HeuristicPool.apply(x_mod, x_map, csr_idx, feat, mode == "max")


# --------------------------------------------------------------------------------------------
# fused feature-map gather + atomic pool (image.py:1285 + pooling.py:63)
# --------------------------------------------------------------------------------------------
def _transpose_last2(t, B, R, S):
    """[B,R,S] -> [B,S,R] copy through dva_transpose_last2 (t contiguous)."""
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        check(_lib.load().dva_transpose_last2(ptr(t), ptr(out), B, R, S, dtype_code(t), stream_ptr()),
              "dva_transpose_last2")
    return out


# NCHW maps: when at least this share of the map's pixels is gathered, one transposition to
# channels-last (2 x map bytes) beats reading every element through its own 32-byte sector
_NCHW_TRANSPOSE_SHARE = 0.25


_INDEX_CHECKS = {"on": os.environ.get("DVA_CHECK_INDICES", "0") not in ("", "0")}


def set_index_checks(on):
    """Validate pixel / image indices of every gather_pool / interp_pool call on the host (one
    device->host read per call) and raise IndexError like the reference's
    `x[feature_map_indexing]` (image.py:1285).  Off by default: the kernels clamp out-of-range
    indices into the map (memory-safe, no synchronisation).  Also switched on by DVA_CHECK_INDICES=1."""
    _INDEX_CHECKS["on"] = bool(on)


def _validate_gather_indices(images, pixels, B, W, H):
    if pixels.numel() == 0:
        return
    lo = torch.stack([pixels[:, 0].min(), pixels[:, 1].min(), images.min()]).tolist()
    hi = torch.stack([pixels[:, 0].max(), pixels[:, 1].max(), images.max()]).tolist()
    if lo[0] < 0 or lo[1] < 0 or lo[2] < 0 or hi[0] >= W or hi[1] >= H or hi[2] >= B:
        raise IndexError(f"mapping out of bounds for feature maps [B={B}, H={H}, W={W}]: pixels x in "
                         f"[{lo[0]}, {hi[0]}], y in [{lo[1]}, {hi[1]}], image ids in [{lo[2]}, {hi[2]}] "
                         f"(stale or mis-scaled mapping?)")


class _GatherPool(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, fmap, images, pixels, atomic_ptr, reduce, channels_last, mapping_size):
        require_cuda(fmap, images, pixels, atomic_ptr)
        lib = _lib.load()
        fmap = fmap.contiguous()
        via_cl = False
        if channels_last:
            B, H, W, C = fmap.shape
        else:
            B, C, H, W = fmap.shape
            n_corner = 1 if mapping_size is None else 4
            if (fmap.dtype in DTYPE_CODES and C % (16 // fmap.element_size()) == 0 and B <= 65535
                    and pixels.shape[0] * n_corner >= _NCHW_TRANSPOSE_SHARE * B * H * W):
                # the reference's layout (image.py:1884): transpose once, then the channels-last kernels
                fmap = _transpose_last2(fmap, B, C, H * W).view(B, H, W, C)
                channels_last, via_cl = True, True
        images = images.long().contiguous()
        if pixels.dtype not in (torch.int16, torch.int32):
            pixels = pixels.int()
        pixels = pixels.contiguous()
        atomic_ptr = _check_csr(atomic_ptr, fmap.device)
        Vw, P, code = atomic_ptr.numel() - 1, pixels.shape[0], REDUCE_CODES[reduce]
        if images.numel() != Vw:
            raise ValueError("images must hold one image id per view (atomic_ptr.numel() - 1)")
        if _INDEX_CHECKS["on"]:
            lim = (W, H) if mapping_size is None else mapping_size
            _validate_gather_indices(images, pixels.long(), B, int(lim[0]), int(lim[1]))
        out = torch.empty((Vw, C), dtype=fmap.dtype, device=fmap.device)
        arg = torch.empty((Vw, C), dtype=torch.int64, device=fmap.device) if code in (2, 3) else None
        head = (ptr(fmap), int(channels_last), ptr(images), ptr(pixels), int(pixels.dtype == torch.int16),
                ptr(atomic_ptr), ptr(out), ptr(arg), B, C, H, W)
        tail = (Vw, P, code, dtype_code(fmap), stream_ptr())
        with torch.cuda.device(fmap.device):
            if mapping_size is None:
                check(lib.dva_gather_pool_fwd(*head, *tail), "dva_gather_pool_fwd")
            else:
                check(lib.dva_interp_pool_fwd(*head, int(mapping_size[0]), int(mapping_size[1]), *tail),
                      "dva_interp_pool_fwd")
        ctx.cfg = (B, C, H, W, Vw, P, code, bool(channels_last), fmap.shape, fmap.dtype, mapping_size, via_cl)
        ctx.save_for_backward(images, pixels, atomic_ptr, arg)
        return out

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        images, pixels, atomic_ptr, arg = ctx.saved_tensors
        B, C, H, W, Vw, P, code, cl, shape, dt, mapping_size, via_cl = ctx.cfg
        lib = _lib.load()
        grad_out = grad_out.contiguous()
        gf = torch.zeros(shape, dtype=torch.float32, device=grad_out.device)
        head = (ptr(grad_out), int(cl), ptr(images), ptr(pixels), int(pixels.dtype == torch.int16),
                ptr(atomic_ptr), ptr(arg), ptr(gf), B, C, H, W)
        tail = (Vw, P, code, dtype_code(grad_out), stream_ptr())
        with torch.cuda.device(grad_out.device):
            if mapping_size is None:
                check(lib.dva_gather_pool_bwd(*head, *tail), "dva_gather_pool_bwd")
            else:
                check(lib.dva_interp_pool_bwd(*head, int(mapping_size[0]), int(mapping_size[1]), *tail),
                      "dva_interp_pool_bwd")
        if via_cl:      # gradient of the NCHW input: transpose the channels-last map gradient back
            gf = _transpose_last2(gf, B, H * W, C).view(B, C, H, W)
        return gf.to(dt), None, None, None, None, None, None


def gather_pool(fmap, images, pixels, atomic_ptr, reduce="max", channels_last=False):
    """segment_csr(fmap[(images_per_pixel, :, py, px)], atomic_ptr, reduce) without the [P,C] copy."""
    return _GatherPool.apply(fmap, images, pixels, atomic_ptr, reduce, channels_last, None)


def interp_pool(fmap, images, pixels, atomic_ptr, mapping_size, reduce="max", channels_last=False):
    """segment_csr(sparse_interpolation(fmap, pixels / (mapping_size - 1), images_per_pixel),
    atomic_ptr, reduce) (image.py:1278-1283 + pooling.py:63) in one kernel.  `pixels` are (x, y)
    at the mapping resolution `mapping_size` = (W_map, H_map); padding mode 'border'."""
    return _GatherPool.apply(fmap, images, pixels, atomic_ptr, reduce, channels_last,
                             (int(mapping_size[0]), int(mapping_size[1])))


def sparse_interpolation_pixels(fmap, images_per_pixel, pixels, mapping_size, channels_last=False):
    """Per-pixel bilinear features [P, C] (image.py:1278-1283): the pooled kernel with one pixel per
    segment."""
    P = pixels.shape[0]
    aptr = torch.arange(P + 1, dtype=torch.int64, device=fmap.device)
    return _GatherPool.apply(fmap, images_per_pixel, pixels, aptr, "sum", channels_last,
                             (int(mapping_size[0]), int(mapping_size[1])))


# --------------------------------------------------------------------------------------------
# fused BatchNorm1d + LeakyReLU over [rows, C] (base_modules.py:38-48, 131-156)
# --------------------------------------------------------------------------------------------
class _BNAct(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, z, weight, bias, running_mean, running_var, training, momentum, eps, slope,
                pre_mean=None, pre_invstd=None):
        require_cuda(z, weight, bias, running_mean, running_var)
        lib = _lib.load()
        z = z.contiguous()
        R, C = z.shape
        dev = z.device
        gamma = weight.detach().float().contiguous() if weight is not None else None
        beta = bias.detach().float().contiguous() if bias is not None else None
        y = torch.empty_like(z)
        # batch statistics already taken in the producing GEMM's epilogue (ops.linear_bn_act): apply only;
        # the backward still differentiates through the batch statistics (ctx keeps training = True)
        have_stats = training and pre_mean is not None
        if have_stats:
            mean, invstd = pre_mean, pre_invstd
        elif training:
            mean = torch.empty(C, dtype=torch.float32, device=dev)
            invstd = torch.empty(C, dtype=torch.float32, device=dev)
        else:
            mean = running_mean.float().contiguous()
            invstd = torch.rsqrt(running_var.float() + eps).contiguous()
        ws_bytes = int(lib.dva_bn_workspace_bytes(R, C))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        rm = running_mean if (training and running_mean is not None) else None
        rv = running_var if (training and running_var is not None) else None
        if not training:
            rm, rv = running_mean, running_var
        kernel_training = training and not have_stats
        if have_stats:
            rm = rv = mean                     # eval-style call: the kernel only reads mean / invstd
        # the training kernel updates the running buffers IN PLACE through raw float pointers
        stage = []
        for name, buf in (("running_mean", rm), ("running_var", rv)):
            if have_stats:
                break
            if buf is not None and (buf.dtype != torch.float32 or not buf.is_contiguous()):
                if not training:
                    raise TypeError(f"{name} must be a contiguous float32 buffer in eval mode")
                stage.append((name, buf, buf.float().contiguous()))   # e.g. a module converted with .half()
        for name, _, tmp in stage:
            if name == "running_mean":
                rm = tmp
            else:
                rv = tmp
        with torch.cuda.device(dev):
            check(lib.dva_bn_act_fwd(ptr(z), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(mean), ptr(invstd), ptr(y),
                                     R, C, float(eps), float(momentum), float(slope), int(bool(kernel_training)),
                                     dtype_code(z), ptr(ws), ws_bytes, stream_ptr()), "dva_bn_act_fwd")
        for _, buf, tmp in stage:
            buf.copy_(tmp)
        ctx.cfg = (R, C, float(slope), bool(training), weight is not None, bias is not None,
                   weight.dtype if weight is not None else None)
        ctx.save_for_backward(z, gamma, beta, mean, invstd)
        return y

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        z, gamma, beta, mean, invstd = ctx.saved_tensors
        R, C, slope, training, has_w, has_b, wdt = ctx.cfg
        lib = _lib.load()
        dy = dy.contiguous()
        dz = torch.empty_like(z)
        sums = torch.empty((2, C), dtype=torch.float32, device=z.device)
        ws_bytes = int(lib.dva_bn_workspace_bytes(R, C))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
        with torch.cuda.device(z.device):
            check(lib.dva_bn_act_bwd(ptr(dy), ptr(z), ptr(gamma), ptr(beta), ptr(mean), ptr(invstd), ptr(dz),
                                     ptr(sums), R, C, slope, int(training), dtype_code(z), ptr(ws), ws_bytes,
                                     stream_ptr()), "dva_bn_act_bwd")
        gw = sums[1].to(wdt) if has_w else None
        gb = sums[0].to(wdt) if has_b else None
        return dz, gw, gb, None, None, None, None, None, None, None, None


def batch_norm_act(z, bn, negative_slope=1.0):
    """act(BatchNorm1d(z)) for z [rows, C] with the statistics / running-average semantics of
    nn.BatchNorm1d (training: batch statistics over all rows, momentum update of the running
    buffers, num_batches_tracked += 1).  `bn` is the nn.BatchNorm1d holding the parameters;
    negative_slope = 1 gives plain BatchNorm, 0.2 the MLP layers of the pools."""
    training = bn.training or (bn.running_mean is None and bn.running_var is None)
    momentum = 0.0 if bn.momentum is None else bn.momentum
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if bn.momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    return _BNAct.apply(z, bn.weight, bn.bias, rm, rv, training, momentum, bn.eps, negative_slope)


# --------------------------------------------------------------------------------------------
# dense projection of the MLP layers on tcgen05 tensor cores (base_modules.py:42)
# --------------------------------------------------------------------------------------------
_GEMM_PRECISION = {"mode": 0}


def set_gemm_precision(mode):
    """Kept for API stability: 'fp32' or 'tf32'.  Every projection kernel is 3xTF32 (fp32-grade
    accuracy) since round 2, so both modes run the same code."""
    _GEMM_PRECISION["mode"] = {"fp32": 0, "tf32": 1}[mode]


def _tc_gemm(a, b, layout, n_out):
    """layout 0: D[M,n_out] = a[M,K] . b[n_out,K]^T;  1: D[M,n_out] = a[M,K] . b[K,n_out];
    2: D[N,n_out] = a[M,N]^T . b[M,n_out]  -- through dva_linear_gemm."""
    lib = _lib.load()
    prec = _GEMM_PRECISION["mode"]
    if layout == 2:
        M, N = a.shape
        K = n_out
        out = torch.empty((N, K), dtype=torch.float32, device=a.device)
    else:
        M, K = a.shape
        N = n_out
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ws_bytes = int(lib.dva_linear_gemm_workspace_bytes(M, N, K, layout, prec))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.dva_linear_gemm(ptr(a), ptr(b), ptr(out), M, N, K, layout, prec, ptr(ws), ws.numel(),
                                  stream_ptr()), "dva_linear_gemm")
    return out


def tc_gemm_supported(x, weight):
    """True for 2-D CUDA floating-point inputs with at least one row: every such projection runs on
    this library's kernels (K, N <= 64: skinny kernels, any K / N; otherwise the tcgen05 kernels, whose
    16-byte TMA rows need K and N to be multiples of 4 -- other widths are zero-padded by `linear`)."""
    return bool(x.is_cuda and weight.is_cuda and x.dim() == 2 and x.shape[0] > 0
                and x.is_floating_point() and weight.is_floating_point())


class _Linear(torch.autograd.Function):
    @staticmethod
    @_fwd_f32
    def forward(ctx, x, weight):
        require_cuda(x, weight)
        x, w = x.float().contiguous(), weight.float().contiguous()
        ctx.save_for_backward(x, w)
        ctx.dtypes = (x.dtype, weight.dtype)
        return _tc_gemm(x, w, 0, w.shape[0])

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, gz):
        x, w = ctx.saved_tensors
        gz = gz.float().contiguous()
        gx = _tc_gemm(gz, w, 1, w.shape[1]) if ctx.needs_input_grad[0] else None
        # dW = dZ^T X: [out,in] result reduced over all rows -- stream-K split over the SMs
        gw = _tc_gemm(gz, x, 2, x.shape[1]) if ctx.needs_input_grad[1] else None
        return gx, gw


class _LinearStats(torch.autograd.Function):
    """z = x @ weight.T with the BatchNorm batch statistics of z's columns taken in the GEMM epilogue
    (dva_linear_bnstats_fwd): returns (z, mean, invstd); running buffers are updated in place."""

    @staticmethod
    @_fwd_f32
    def forward(ctx, x, weight, running_mean, running_var, momentum, eps):
        require_cuda(x, weight)
        lib = _lib.load()
        x, w = x.float().contiguous(), weight.float().contiguous()
        M, K = x.shape
        N = w.shape[0]
        z = torch.empty((M, N), dtype=torch.float32, device=x.device)
        mean = torch.empty(N, dtype=torch.float32, device=x.device)
        invstd = torch.empty(N, dtype=torch.float32, device=x.device)
        stage = []
        rm, rv = running_mean, running_var
        for name, buf in (("m", rm), ("v", rv)):
            if buf is not None and (buf.dtype != torch.float32 or not buf.is_contiguous()):
                stage.append((name, buf, buf.float().contiguous()))
        for name, _, tmp in stage:
            if name == "m":
                rm = tmp
            else:
                rv = tmp
        ws_bytes = int(lib.dva_linear_bnstats_workspace_bytes(N, K))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.dva_linear_bnstats_fwd(ptr(x), ptr(w), ptr(z), M, N, K, float(eps), float(momentum), ptr(mean),
                                             ptr(invstd), ptr(rm), ptr(rv), ptr(ws), ws_bytes, stream_ptr()),
                  "dva_linear_bnstats_fwd")
        for _, buf, tmp in stage:
            buf.copy_(tmp)
        ctx.save_for_backward(x, w)
        ctx.mark_non_differentiable(mean, invstd)
        return z, mean, invstd

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, gz, _gm, _gi):
        x, w = ctx.saved_tensors
        gz = gz.float().contiguous()
        gx = _tc_gemm(gz, w, 1, w.shape[1]) if ctx.needs_input_grad[0] else None
        gw = _tc_gemm(gz, x, 2, x.shape[1]) if ctx.needs_input_grad[1] else None
        return gx, gw, None, None, None, None


class _MLPLayer(torch.autograd.Function):
    """One narrow MLP layer act(BatchNorm1d(x @ weight.T)) in training mode (base_modules.py:38-48) as ONE autograd
    node: forward = GEMM with the batch statistics in its epilogue (dva_linear_bnstats_fwd) + the apply pass;
    backward = the statistics pass + ONE kernel for dz (kept on chip), dX and dW (dva_mlp_layer_bwd)."""

    @staticmethod
    @_fwd_f32
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, momentum, eps, slope):
        require_cuda(x, weight, gamma, beta)
        lib = _lib.load()
        x, w = x.float().contiguous(), weight.float().contiguous()
        M, K = x.shape
        N = w.shape[0]
        dev = x.device
        z = torch.empty((M, N), dtype=torch.float32, device=dev)
        mean = torch.empty(N, dtype=torch.float32, device=dev)
        invstd = torch.empty(N, dtype=torch.float32, device=dev)
        g = gamma.detach().float().contiguous() if gamma is not None else None
        b = beta.detach().float().contiguous() if beta is not None else None
        stage = []
        rm, rv = running_mean, running_var
        for name, buf in (("m", rm), ("v", rv)):
            if buf is not None and (buf.dtype != torch.float32 or not buf.is_contiguous()):
                stage.append((name, buf, buf.float().contiguous()))
        for name, _, tmp in stage:
            if name == "m":
                rm = tmp
            else:
                rv = tmp
        ws_bytes = max(int(lib.dva_linear_bnstats_workspace_bytes(N, K)), int(lib.dva_bn_workspace_bytes(M, N)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        y = torch.empty_like(z)
        with torch.cuda.device(dev):
            check(lib.dva_linear_bnstats_fwd(ptr(x), ptr(w), ptr(z), M, N, K, float(eps), float(momentum), ptr(mean),
                                             ptr(invstd), ptr(rm), ptr(rv), ptr(ws), ws_bytes, stream_ptr()),
                  "dva_linear_bnstats_fwd")
            # apply half: eval-style call on the batch statistics (the kernel only reads mean / invstd)
            check(lib.dva_bn_act_fwd(ptr(z), ptr(g), ptr(b), ptr(mean), ptr(mean), ptr(mean), ptr(invstd), ptr(y),
                                     M, N, float(eps), 0.0, float(slope), 0, dtype_code(z), ptr(ws), ws_bytes,
                                     stream_ptr()), "dva_bn_act_fwd")
        for _, buf, tmp in stage:
            buf.copy_(tmp)
        ctx.cfg = (float(slope), gamma is not None, beta is not None,
                   gamma.dtype if gamma is not None else (beta.dtype if beta is not None else None), weight.dtype)
        ctx.save_for_backward(x, w, z, g, b, mean, invstd)
        return y

    @staticmethod
    @_bwd
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, w, z, g, b, mean, invstd = ctx.saved_tensors
        slope, has_g, has_b, pdt, wdt = ctx.cfg
        lib = _lib.load()
        M, K = x.shape
        N = w.shape[0]
        dy = dy.float().contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        sums = torch.empty((2, N), dtype=torch.float32, device=x.device)
        ws_bytes = int(lib.dva_mlp_layer_bwd_workspace_bytes(M, N, K))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.dva_mlp_layer_bwd(ptr(dy), ptr(z), ptr(x), ptr(w), ptr(g), ptr(b), ptr(mean), ptr(invstd),
                                        ptr(dx), ptr(dw), ptr(sums), M, N, K, slope, ptr(ws), ws_bytes, stream_ptr()),
                  "dva_mlp_layer_bwd")
        gw = sums[1].to(pdt) if has_g else None
        gb = sums[0].to(pdt) if has_b else None
        return dx, (dw.to(wdt) if ctx.needs_input_grad[1] else None), gw, gb, None, None, None, None, None


# The fused backward (3xTF32 on mma.sync) takes 0.106 ms in a step at 1.28 M x 32 x 32 (HBM time: 0.100) against
# 0.26 - 0.34 ms for the three kernels it replaces, but is no faster than them at K = 64 (0.33 against 0.29 ms),
# where dX rides the tcgen05 kernel -> layers with K <= 32 only.
_MLP_LAYER_FUSED = {"on": os.environ.get("DVA_MLP_LAYER_FUSED", "1") != "0", "max_k": 32}


def linear_bn_act(x, weight, bn, negative_slope=1.0):
    """act(BatchNorm1d(x @ weight.T)): one MLP layer of the pools (base_modules.py:38-48).  In training,
    when the layer is wide enough for the tcgen05 kernel and has at most 128 output channels, the batch
    statistics come out of the GEMM epilogue (2 passes over the activations instead of 3); otherwise
    linear() followed by batch_norm_act()."""
    lib = _lib.load()
    M, K = x.shape
    N = weight.shape[0]
    training = bn.training or (bn.running_mean is None and bn.running_var is None)
    if not (training and x.is_cuda and M > 0 and K % 4 == 0 and lib.dva_linear_bnstats_supported(M, N, K)):
        return batch_norm_act(linear(x, weight), bn, negative_slope=negative_slope)
    momentum = 0.0 if bn.momentum is None else bn.momentum
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if bn.momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    if (_MLP_LAYER_FUSED["on"] and K <= _MLP_LAYER_FUSED["max_k"] and lib.dva_mlp_layer_bwd_supported(M, N, K)
            and torch.is_grad_enabled()
            and (x.requires_grad or weight.requires_grad)):
        return _MLPLayer.apply(x, weight, bn.weight, bn.bias, rm, rv, momentum, bn.eps, negative_slope)
    z, mean, invstd = _LinearStats.apply(x, weight, rm, rv, momentum, bn.eps)
    return _BNAct.apply(z, bn.weight, bn.bias, rm, rv, True, momentum, bn.eps, negative_slope, mean, invstd)


def linear(x, weight):
    """x @ weight.T for a bias-free nn.Linear weight [out, in] (base_modules.py:42), always on this
    library's kernels, computed from fp32 operands (also under autocast).  Wide layers whose K or N is
    not a multiple of 4 are zero-padded to the next multiple (exact: the padding contributes 0)."""
    if not tc_gemm_supported(x, weight):
        raise RuntimeError("ops.linear needs 2-D CUDA floating-point operands with at least one row "
                           "(no CPU / library fallback)")
    out_dtype = x.dtype if not torch.is_autocast_enabled("cuda") else torch.float32
    K, N = x.shape[1], weight.shape[0]
    if not (K <= 64 and N <= 64):
        pk, pn = (-K) % 4, (-N) % 4
        if pk:
            x = torch.nn.functional.pad(x, (0, pk))
            weight = torch.nn.functional.pad(weight, (0, pk))
        if pn:
            weight = torch.nn.functional.pad(weight, (0, 0, 0, pn))
        z = _Linear.apply(x, weight)
        z = z[:, :N] if pn else z
    else:
        z = _Linear.apply(x, weight)
    return z if z.dtype == out_dtype or out_dtype not in (torch.float16, torch.bfloat16) else z.to(out_dtype)
