"""config #0 (toy: 1k points x ~4 views through the DeepViewAgg module): our UnimodalBranch on the
GPU vs the reference's UnimodalBranch executed on CPU (tests/golden/unimodal_branch_toy.npz):
output features, seen mask, gradients w.r.t. the 3D features, the 2D feature maps and every
view-pool parameter.  Also the container kernels on CUDA tensors."""
import pytest
import torch

from conftest import load_golden
from test_containers import _toy_image_data, canon_pixels

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
    assert err <= tol, f"{what}: rel err {err:.3e} > {tol}"


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("interpolate", [False, True])
def test_unimodal_branch_vs_reference(channels_last, interpolate):
    """interpolate=True: setting 0 (maps at half resolution) goes through the fused bilinear
    interpolation + max-pool kernel (image.py:1278-1283)."""
    from deepviewagg_b200.modules.multimodal.fusion import BimodalFusion
    from deepviewagg_b200.modules.multimodal.modules import UnimodalBranch
    from deepviewagg_b200.modules.multimodal.pooling import BimodalCSRPool, GroupBimodalCSRPool
    g = load_golden("unimodal_branch_interp" if interpolate else "unimodal_branch_toy")
    mod = _toy_image_data(g, "cuda")
    xs = []
    for im in mod:
        x = im.x.detach().clone()
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        x.requires_grad_(True)
        im._x = x
        xs.append(x)
    view_pool = GroupBimodalCSRPool(in_map=8, in_mod=16, num_groups=4, use_num=True)
    view_pool.load_state_dict(g["sd"], strict=True)
    branch = UnimodalBranch(None, BimodalCSRPool(mode="max"), view_pool, BimodalFusion("concatenation"),
                            interpolate=interpolate).cuda()
    branch.train()
    x_3d = g["x_3d"].cuda().requires_grad_(True)
    out = branch({"x_3d": x_3d, "x_seen": None, "modalities": {"image": mod}}, "image")
    assert branch.out_channels == 28
    _close(out["x_3d"], g["out"], 1e-4, "x_3d out")
    assert torch.equal(out["x_seen"].cpu(), g["x_seen"])
    params = dict(view_pool.named_parameters())
    grads = torch.autograd.grad((out["x_3d"] * g["w"].cuda()).sum(), [x_3d] + xs + list(params.values()),
                                allow_unused=True)
    names = ["x_3d", "s0_x", "s1_x"] + ["param/" + k for k in params]
    for n, gr in zip(names, grads):
        ref = g["grad"][n]
        if gr is None:
            assert float(ref.abs().max()) == 0, n
            continue
        assert (gr.cpu() - ref).abs().max() <= 2e-4 * max(1.0, float(ref.abs().max())), n


def test_branch_empty_modality_and_identity():
    from deepviewagg_b200.modules.multimodal.fusion import BimodalFusion
    from deepviewagg_b200.modules.multimodal.modules import IdentityBranch, MultimodalBlockDown, UnimodalBranch
    from deepviewagg_b200.modules.multimodal.pooling import BimodalCSRPool
    from deepviewagg_b200.core.multimodal.image import ImageData
    branch = UnimodalBranch(None, BimodalCSRPool("max"), BimodalCSRPool("mean"), BimodalFusion("concatenation"),
                            out_channels=20).cuda()
    x_3d = torch.randn(7, 12, device="cuda")
    d = branch({"x_3d": x_3d, "x_seen": None, "modalities": {"image": ImageData([])}}, "image")
    assert d["x_3d"].shape == (7, 20) and (d["x_3d"][:, 12:] == 0).all() and d["x_seen"] is None
    blk = MultimodalBlockDown(None, None, image=IdentityBranch())
    d2 = {"x_3d": x_3d, "x_seen": None, "modalities": {}}
    assert blk(d2) is d2


def test_containers_on_cuda_match_cpu():
    from deepviewagg_b200.core.multimodal.image import ImageMapping
    g = load_golden("image_mapping")
    args = [g[k] for k in ("point_ids", "image_ids", "pixels", "features")]
    m_cpu = ImageMapping.from_dense(*args, num_points=int(g["num_points"]))
    m = ImageMapping.from_dense(*[a.cuda() for a in args], num_points=int(g["num_points"]))
    assert torch.equal(m.pointers.cpu(), g["pointers"]) and torch.equal(m.images.cpu(), g["images"])
    assert torch.equal(m.atomic_csr_indexing.cpu(), g["atomic_pointers"])
    assert torch.allclose(m.features.cpu(), g["out_features"], atol=1e-6)
    ms = m.select_points(g["sel"].cuda(), mode="pick")
    assert torch.equal(ms.pointers.cpu(), g["sel_pointers"]) and torch.equal(ms.images.cpu(), g["sel_images"])
    gm = load_golden("image_mapping_merge")
    mg = m.select_points(gm["merge_idx"].cuda(), mode="merge")
    assert torch.equal(mg.pointers.cpu(), gm["pointers"]) and torch.equal(mg.images.cpu(), gm["images"])
    assert torch.equal(canon_pixels(mg.pixels.cpu(), mg.atomic_csr_indexing.cpu()),
                       canon_pixels(gm["pixels"], gm["atomic_pointers"]))
    assert torch.allclose(mg.features.cpu(), gm["features"], atol=1e-6)
    assert torch.equal(m_cpu.pixels, m.pixels.cpu())


@pytest.mark.parametrize("tag", ["half", "quarter"])
@pytest.mark.parametrize("channels_last", [False, True])
def test_sparse_interpolation_kernel_bit_exact(tag, channels_last):
    """dva_interp_pool_fwd/bwd vs the reference's sparse_interpolation executed on CPU: fp32 values
    bit-identical (same operation order), map gradient within fp32 atomics' reordering; then pooled
    (max / mean / sum over random pixel segments) against the oracle composition."""
    import numpy as np
    from deepviewagg_b200 import ops
    from oracle.image_oracle import sparse_interpolation_pixels
    from oracle import scatter_standin as S
    g = load_golden("sparse_interpolation")
    W, H, ds = [int(v) for v in g[f"{tag}_size"]]
    x = g[f"{tag}_x"].cuda()
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    fmap = x.permute(0, 2, 3, 1) if channels_last else x
    pix, batch = g[f"{tag}_pix"].cuda(), g[f"{tag}_batch"].cuda()
    for pdt in (torch.int16, torch.int32, torch.int64):
        out = ops.sparse_interpolation_pixels(fmap, batch, pix.to(pdt), (W, H), channels_last=channels_last)
        assert torch.equal(out.cpu(), g[f"{tag}_out"]), pdt
    (gx,) = torch.autograd.grad((out * g[f"{tag}_w"].cuda()).sum(), [x])
    assert (gx.cpu() - g[f"{tag}_gx"]).abs().max() <= 1e-5 * float(g[f"{tag}_gx"].abs().max())
    # pooled: pixels sorted by image so that a segment lives in one image, ragged segments with empties
    order = torch.argsort(batch.cpu(), stable=True)
    pix_s, batch_s = pix.cpu()[order], batch.cpu()[order]
    gen = torch.Generator().manual_seed(5)
    ptr = [0]
    for b in range(int(batch_s.max()) + 1):
        lo, hi = int((batch_s < b).sum()), int((batch_s <= b).sum())
        cuts = torch.sort(torch.randint(lo, hi + 1, (40,), generator=gen)).values.tolist()
        ptr += cuts + [hi]
    ptr = torch.tensor(ptr)
    img = batch_s[torch.clamp(ptr[:-1], max=batch_s.numel() - 1)]
    vals = torch.from_numpy(sparse_interpolation_pixels(g[f"{tag}_x"].numpy(), pix_s.numpy(), batch_s.numpy(), (W, H)))
    for reduce in ("max", "min", "sum", "mean"):
        got = ops.interp_pool(fmap.detach(), img.cuda(), pix_s.int().cuda(), ptr.cuda(), (W, H), reduce=reduce,
                              channels_last=channels_last)
        want = S.segment_csr(vals, ptr, reduce=reduce)
        if reduce in ("max", "min"):
            assert torch.equal(got.cpu(), want), reduce
        else:
            assert (got.cpu() - want).abs().max() <= 1e-5 * max(1.0, float(want.abs().max())), reduce


def test_flat_mapping_file_uploads_to_cuda(tmp_path):
    """storage.load_image_data(device='cuda'): memmap -> pinned staging -> async H2D, tensors equal
    the originals and the loaded ImageData drives the fused branch kernels."""
    from deepviewagg_b200.core.multimodal.storage import load_image_data, save_image_data
    g = load_golden("unimodal_branch_toy")
    mod = _toy_image_data(g)
    path = save_image_data(str(tmp_path / "scene.dvamap"), mod)
    back = load_image_data(path, device="cuda")
    torch.cuda.synchronize()
    for a, b in zip(mod, back):
        assert b.mappings.pointers.is_cuda and b.mappings.pixels.is_cuda
        assert torch.equal(a.mappings.pointers, b.mappings.pointers.cpu())
        assert torch.equal(a.mappings.pixels, b.mappings.pixels.cpu())
        assert torch.equal(a.mappings.features, b.mappings.features.cpu())
    assert torch.equal(back.view_cat_csr_indexing.cpu(), mod.view_cat_csr_indexing)
