"""U2: MultimodalBlockDown.forward_3d_block_down on NON-Identity blocks against the executed reference
(modules.py:101-236; fixture tests/golden/block_down.npz written by oracle/make_golden.py):
  'pick'  -- dense block with a sampler: x_seen[idx] and every setting's mappings re-indexed
  'merge' -- strided sparse block (fake MinkowskiEngine tensor): child->parent index from the
             coordinate map, x_seen OR-ed over children (:225), select_points(idx, 'merge')
             (image.py:2211-2273).
Integers bit-exact after canonicalising the implementation-defined order of pixels within a view.
Runs on CPU tensors (host logic) and, marked gpu, on CUDA tensors (the kernels' path)."""
import types

import pytest
import torch

from conftest import load_golden
from deepviewagg_b200.core.multimodal.image import ImageData, ImageMapping, SameSettingImageData
from deepviewagg_b200.modules.multimodal import modules as M
from test_containers import canon_pixels


class _Sampler:
    last_idx = None


class PickBlock(torch.nn.Module):
    def __init__(self, idx):
        super().__init__()
        self.sampler, self.idx = _Sampler(), idx

    def forward(self, x):
        self.sampler.last_idx = self.idx
        return x[self.idx]


class FakeCoordsManager:
    def __init__(self, src, target):
        self.src, self.target = src, target

    def get_coords_map(self, stride_in, stride_out):
        return self.src, self.target


class FakeSparseTensor:
    def __init__(self, F, stride, coords_man):
        self.F, self.tensor_stride, self.coords_man = F, [stride], coords_man


class StridedBlock(torch.nn.Module):
    def __init__(self, parent):
        super().__init__()
        self.parent = parent

    def forward(self, x):
        n_out = int(self.parent.max()) + 1
        F = torch.zeros(n_out, x.F.shape[1], device=x.F.device).index_add_(0, self.parent, x.F)
        cnt = torch.zeros(n_out, device=x.F.device).index_add_(0, self.parent, torch.ones_like(self.parent, dtype=torch.float))
        return FakeSparseTensor(F / cnt.clamp(min=1).unsqueeze(1), x.tensor_stride[0] * 2, x.coords_man)


def _image_data(g, device):
    ims = []
    for s in (0, 1):
        W, H, n_img = [int(v) for v in g[f"s{s}_size"]]
        im = SameSettingImageData(pos=torch.zeros(n_img, 3), opk=torch.zeros(n_img, 3), ref_size=(W, H),
                                  proj_upscale=1, downscale=1)
        im.mappings = ImageMapping.from_dense(g[f"s{s}_pid"], g[f"s{s}_iid"], g[f"s{s}_pix"], g[f"s{s}_feat"],
                                              num_points=int(g["n_points"]))
        ims.append(im)
    return ImageData(ims).to(device)


def _check_mappings(mod, g, prefix):
    for s, im in enumerate(mod):
        m = im.mappings
        assert int(im.num_views) == int(g[f"{prefix}s{s}_num_views"])
        assert torch.equal(m.pointers.cpu(), g[f"{prefix}s{s}_pointers"]), (prefix, s, "pointers")
        assert torch.equal(m.images.cpu(), g[f"{prefix}s{s}_images"]), (prefix, s, "images")
        ap = m.values[1].pointers.cpu()
        assert torch.equal(ap, g[f"{prefix}s{s}_atomic_pointers"]), (prefix, s, "atomic pointers")
        assert m.pixels.dtype == g[f"{prefix}s{s}_pixels"].dtype
        assert torch.equal(canon_pixels(m.pixels.cpu(), ap), canon_pixels(g[f"{prefix}s{s}_pixels"], ap))
        assert torch.allclose(m.features.cpu(), g[f"{prefix}s{s}_features"], rtol=1e-6, atol=1e-7)


def _run(device):
    g = load_golden("block_down")
    x_3d, x_seen = g["x_3d"].to(device), g["x_seen"].to(device)
    # (a) pick
    d = M.MultimodalBlockDown.forward_3d_block_down(
        {"x_3d": x_3d.clone(), "x_seen": x_seen.clone(), "modalities": {"image": _image_data(g, device)}},
        PickBlock(g["pick_idx"].to(device)))
    assert torch.equal(d["x_3d"].cpu(), g["pick_x_3d"]) and torch.equal(d["x_seen"].cpu(), g["pick_x_seen"])
    _check_mappings(d["modalities"]["image"], g, "pick_")
    # (b) merge through the MinkowskiEngine branch
    parent, src = g["merge_parent"].to(device), g["merge_src"].to(device)
    cm = FakeCoordsManager(src, parent[src])
    saved = M.me
    M.me = types.SimpleNamespace(SparseTensor=FakeSparseTensor)
    try:
        d = M.MultimodalBlockDown.forward_3d_block_down(
            {"x_3d": FakeSparseTensor(x_3d.clone(), 1, cm), "x_seen": x_seen.clone(),
             "modalities": {"image": _image_data(g, device)}}, StridedBlock(parent))
    finally:
        M.me = saved
    assert torch.allclose(d["x_3d"].F.cpu(), g["merge_x_3d"], atol=1e-6)
    assert d["x_seen"].dtype == torch.bool and torch.equal(d["x_seen"].cpu(), g["merge_x_seen"])
    _check_mappings(d["modalities"]["image"], g, "merge_")
    # a block that samples nothing (identity permutation) leaves the mappings untouched (:137-140)
    mod = _image_data(g, device)
    keep = [im.mappings.pointers.clone() for im in mod]
    d = M.MultimodalBlockDown.forward_3d_block_down(
        {"x_3d": x_3d.clone(), "x_seen": x_seen.clone(), "modalities": {"image": mod}},
        PickBlock(torch.arange(x_3d.shape[0], device=device)))
    assert torch.equal(d["x_seen"], x_seen)
    assert all(torch.equal(im.mappings.pointers, k) for im, k in zip(d["modalities"]["image"], keep))


def test_forward_3d_block_down_cpu_tensors():
    _run("cpu")


@pytest.mark.gpu
def test_forward_3d_block_down_cuda_tensors():
    _run("cuda")
