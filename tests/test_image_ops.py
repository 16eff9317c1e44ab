"""I3 / I4: the remaining ImageMapping re-indexing operations against the executed reference
(fixture tests/golden/image_ops.npz from oracle/make_golden.py): select_images (image.py:2029-2093),
select_views (:2095-2165), crop (:2279-2342), downscale_images / upscale_images (:1916-2027).
Integers bit-exact (pixels compared after canonicalising their implementation-defined order within a view).
Runs on CPU tensors and, marked gpu, on CUDA tensors."""
import pytest
import torch

from conftest import load_golden
from deepviewagg_b200.core.multimodal.image import ImageMapping
from test_containers import canon_pixels


def _check(mm, g, tag):
    assert torch.equal(mm.pointers.cpu(), g[f"{tag}_pointers"]), (tag, "pointers")
    assert torch.equal(mm.images.cpu(), g[f"{tag}_images"]), (tag, "images")
    ap = mm.values[1].pointers.cpu()
    assert torch.equal(ap, g[f"{tag}_atomic_pointers"]), (tag, "atomic pointers")
    assert mm.pixels.dtype == g[f"{tag}_pixels"].dtype, (tag, mm.pixels.dtype)
    assert torch.equal(canon_pixels(mm.pixels.cpu(), ap), canon_pixels(g[f"{tag}_pixels"], ap)), (tag, "pixels")
    assert torch.allclose(mm.features.cpu(), g[f"{tag}_features"], rtol=1e-6, atol=1e-7), (tag, "features")


def _run(device):
    g = load_golden("image_ops")
    m = ImageMapping.from_dense(g["point_ids"].to(device), g["image_ids"].to(device), g["pixels"].to(device),
                                g["features"].to(device), num_points=int(g["num_points"]))
    _check(m.select_images(g["img_idx"].to(device)), g, "select_images")
    mv, seen = m.select_views(g["view_mask"].to(device))
    _check(mv, g, "select_views")
    assert torch.equal(seen.cpu(), g["select_views_img_idx"])
    _check(m.crop(tuple(int(v) for v in g["crop_size"]), g["crop_offsets"].to(device)), g, "crop")
    _check(m.downscale_images(4), g, "down4")
    _check(m.upscale_images(2), g, "up2")
    _check(m.upscale_images(2, center=False), g, "up2_nocenter")


def test_image_ops_cpu_tensors():
    _run("cpu")


@pytest.mark.gpu
def test_image_ops_cuda_tensors():
    _run("cuda")
