"""Host logic of the mapping containers (CPU tensors, no kernel launches): CSRData / CSRBatch /
ImageMapping / ImageData mirror vs fixtures of the executed reference and the reference's own
round-trip script (image.py:2350-2390).  Integers bit-exact after canonicalising unstable-sort ties."""
import numpy as np
import torch

from conftest import load_golden
from deepviewagg_b200.core.multimodal.csr import CSRBatch, CSRData
from deepviewagg_b200.core.multimodal.image import (ImageBatch, ImageData, ImageMapping, ImageMappingBatch,
                                                    SameSettingImageData)
from deepviewagg_b200.utils.multimodal import lexargsort, lexargunique, lexsort, lexunique


def canon_pixels(pixels, atomic_ptr):
    """sort the pixels of every view by (x, y): their order is implementation-defined."""
    out = pixels.clone().long()
    ap = atomic_ptr.tolist()
    for a, b in zip(ap[:-1], ap[1:]):
        if b - a > 1:
            seg = out[a:b]
            out[a:b] = seg[torch.sort(seg[:, 0] * 100000 + seg[:, 1]).indices]
    return out


def test_lex_ops_kat():
    g = load_golden("kat_lex")
    a, b = g["a"], g["b"]
    s = lexargsort(a, b)
    assert (a[s].tolist(), b[s].tolist()) == ([0, 0, 1, 2, 2], [5, 5, 3, 0, 1])
    assert lexargunique(a, b).tolist() == g["argunique"].tolist() == [1, 3, 2, 0]
    u = lexunique(a, b)
    assert u[0].tolist() == g["unique_a"].tolist() and u[1].tolist() == g["unique_b"].tolist()
    sa, sb = lexsort(a, b)
    assert sa.tolist() == [0, 0, 1, 2, 2] and sb.tolist() == [5, 5, 3, 0, 1]
    # stability: equal keys keep their input order
    assert lexargsort(torch.tensor([1, 1, 0, 1])).tolist() == [2, 0, 1, 3]


def test_image_mapping_matches_reference():
    g = load_golden("image_mapping")
    m = ImageMapping.from_dense(g["point_ids"], g["image_ids"], g["pixels"], g["features"],
                                num_points=int(g["num_points"]))
    assert torch.equal(m.pointers, g["pointers"]) and torch.equal(m.images, g["images"])
    assert torch.equal(m.atomic_csr_indexing, g["atomic_pointers"])
    assert m.pixels.dtype == torch.int16 and m.pointers.dtype == torch.int64
    assert torch.equal(canon_pixels(m.pixels, m.atomic_csr_indexing),
                       canon_pixels(g["out_pixels"], g["atomic_pointers"]))
    assert torch.allclose(m.features, g["out_features"], atol=1e-6)
    fmi = m.feature_map_indexing
    assert torch.equal(fmi[0], g["fmi_batch"]) and fmi[1] is Ellipsis
    # 'pick' selection (csr.py:266-294)
    ms = m.select_points(g["sel"], mode="pick")
    assert torch.equal(ms.pointers, g["sel_pointers"]) and torch.equal(ms.images, g["sel_images"])
    assert torch.equal(ms.atomic_csr_indexing, g["sel_atomic_pointers"])
    assert torch.allclose(ms.features, g["sel_features"], atol=1e-6)
    # downscale: pix // 4; the reference never drops duplicates (its dedupe key is the item id)
    d = m.downscale_images(4)
    assert torch.equal(d.atomic_csr_indexing, g["down_atomic_pointers"])
    assert torch.equal(canon_pixels(d.pixels, d.atomic_csr_indexing),
                       canon_pixels(g["down_pixels"], g["down_atomic_pointers"]))
    assert torch.equal(m.upscale_images(2).pixels, (m.pixels.float() * 2 + 1).long().to(m.pixels.dtype))
    # 'merge' after a strided 3D conv (image.py:2211-2273)
    gm = load_golden("image_mapping_merge")
    mg = m.select_points(gm["merge_idx"], mode="merge")
    assert torch.equal(mg.pointers, gm["pointers"]) and torch.equal(mg.images, gm["images"])
    assert torch.equal(mg.atomic_csr_indexing, gm["atomic_pointers"])
    assert torch.equal(canon_pixels(mg.pixels, mg.atomic_csr_indexing),
                       canon_pixels(gm["pixels"], gm["atomic_pointers"]))
    assert torch.allclose(mg.features, gm["features"], atol=1e-6)


def test_reference_round_trip_script():
    """image.py:2350-2390: the reference prints True x3 for this script."""
    gen = torch.Generator().manual_seed(0)
    n_groups, n_items = 1000, 10000
    idx = torch.randint(0, n_groups, (n_items,), generator=gen)
    img_idx = torch.randint(0, 3, (n_items,), generator=gen)
    pixels = torch.randint(0, 10, (n_items, 2), generator=gen)
    features = torch.rand(n_items, 3, generator=gen)
    idx, img_idx = lexsort(idx, img_idx)
    m = ImageMapping.from_dense(idx, img_idx, pixels, features)
    b = ImageMappingBatch.from_csr_list([m[2], m[1:3], m, m[0]])
    assert isinstance(b, ImageMappingBatch) and b.num_batch_items == 4
    a = m[2].num_groups + m[1:3].num_groups
    assert (b[a:a + m.num_groups].values[1].values[0] == m.values[1].values[0]).all()
    back = b.to_csr_list()
    assert (back[2].pointers == m.pointers).all()
    assert (back[2].values[1].values[0] == m.values[1].values[0]).all()
    assert torch.equal(back[2].images, m.images)
    # plain CSR batch indexing (second half of the script)
    c = CSRData(torch.tensor([0, 0, 5, 12, 12, 15]), torch.arange(15), dense=False)
    cb = CSRBatch.from_csr_list([c, c, c])
    sel = cb[[0, 0, 5]]
    assert sel.pointers.tolist() == [0, 0, 0, 0] and sel.num_items == 0
    sel = cb[[1, 7, 14]]
    assert sel.pointers.tolist() == [0, 5, 12, 15] and sel.values[0].tolist() == (
        list(range(0, 5)) + list(range(5, 12)) + list(range(12, 15)))


def test_csr_empty_groups_and_reindex():
    c = CSRData(torch.tensor([0, 0, 1, 1, 3]), torch.tensor([10, 11, 12, 13, 14]), dense=True)
    assert c.pointers.tolist() == [0, 2, 4, 5]
    c.insert_empty_groups(torch.tensor([1, 2, 5]), num_groups=8)
    assert c.pointers.tolist() == [0, 0, 2, 4, 4, 4, 5, 5, 5]
    r = CSRData(torch.tensor([0, 2, 3]), torch.tensor([7, 8, 9])).reindex_groups(torch.tensor([3, 0]))
    assert r.pointers.tolist() == [0, 1, 1, 1, 3] and r.values[0].tolist() == [9, 7, 8]


def _toy_image_data(g, device="cpu"):
    ims = []
    for s in (0, 1):
        W, H, n_img = [int(v) for v in g[f"s{s}_size"]]
        im = SameSettingImageData(pos=torch.zeros(n_img, 3), opk=torch.zeros(n_img, 3), ref_size=(W, H),
                                  proj_upscale=1, downscale=1)
        im.mappings = ImageMapping.from_dense(g[f"s{s}_pid"], g[f"s{s}_iid"], g[f"s{s}_pix"], g[f"s{s}_feat"],
                                              num_points=int(g["n_points"]))
        im.x = g[f"s{s}_x"].clone()
        ims.append(im)
    return ImageData(ims).to(device)


def test_image_data_view_indexing_matches_reference():
    g = load_golden("unimodal_branch_toy")
    mod = _toy_image_data(g)
    assert mod[0].downscale == 2 and mod[1].downscale == 1          # x setter updates the scale
    assert torch.equal(mod.view_cat_csr_indexing, g["csr"])
    srt = mod.view_cat_sorting
    dense = torch.cat([torch.arange(im.num_points).repeat_interleave(
        im.view_csr_indexing[1:] - im.view_csr_indexing[:-1]) for im in mod])
    assert (dense[srt][1:] >= dense[srt][:-1]).all() and srt.unique().numel() == srt.numel()
    # mapped features at the feature-map resolution: [P, C] with P = number of views here
    feats = mod.get_mapped_features(interpolate=False)
    assert feats[0].shape == (mod[0].mappings.num_items, 16)
    assert torch.equal(feats[0], mod[0].x[mod[0].scaled_mappings().feature_map_indexing])
    # batching two copies: points and images are offset, pointers concatenated
    b = ImageBatch.from_data_list([mod, mod.clone()])
    assert b.num_points == 2 * mod.num_points and b[0].num_views == 2 * mod[0].num_views
    n = mod.num_points
    assert torch.equal(b[0].mappings.pointers[:n + 1], mod[0].mappings.pointers)
    assert int(b[0].mappings.images.max()) == 2 * mod[0].num_views - 1
    # selecting points drops unseen images and renumbers the rest
    sub = mod.select_points(torch.arange(0, 50), mode="pick")
    assert sub.num_points == 50 and all(int(im.mappings.images.max()) < im.num_views for im in sub)


def test_flat_mapping_file_round_trip(tmp_path):
    """the flat on-disk mapping format reloads to identical tensors (memmap
    read, no pickling) and keeps the container behaviour (view sorting, point selection)."""
    from deepviewagg_b200.core.multimodal.storage import load_image_data, read_header, save_image_data
    g = load_golden("unimodal_branch_toy")
    mod = _toy_image_data(g)
    mod[0].extras["extrinsic"] = torch.arange(mod[0].num_views * 16, dtype=torch.float64).view(-1, 4, 4)
    path = save_image_data(str(tmp_path / "scene.dvamap"), mod)
    meta, base = read_header(path)
    assert base % 64 == 0 and all(e["offset"] % 64 == 0 for e in meta["arrays"].values())
    back = load_image_data(path)
    assert len(back) == len(mod)
    for a, b in zip(mod, back):
        assert torch.equal(a.mappings.pointers, b.mappings.pointers)
        assert torch.equal(a.mappings.images, b.mappings.images)
        assert torch.equal(a.mappings.values[1].pointers, b.mappings.values[1].pointers)
        assert torch.equal(a.mappings.pixels, b.mappings.pixels) and a.mappings.pixels.dtype == b.mappings.pixels.dtype
        assert torch.equal(a.mappings.features, b.mappings.features)
        assert a.ref_size == b.ref_size and a.downscale == b.downscale and a.num_views == b.num_views
        assert torch.equal(a.pos, b.pos)
    assert torch.equal(back[0].extras["extrinsic"], mod[0].extras["extrinsic"])
    assert torch.equal(back.view_cat_csr_indexing, mod.view_cat_csr_indexing)
    idx = torch.arange(0, mod.num_points, 3)
    sa, sb = mod.select_points(idx), back.select_points(idx)
    assert torch.equal(sa[0].mappings.pointers, sb[0].mappings.pointers)
    assert torch.equal(sa[1].mappings.images, sb[1].mappings.images)
