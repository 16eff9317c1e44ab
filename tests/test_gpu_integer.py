"""GPU parity of the integer kernels (projection / splat boxes / z-buffer / CSR bookkeeping)
against fixtures of the executed numba reference and against the C oracle.  Bit-exact."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import visibility_oracle as VO

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


@pytest.mark.parametrize("tag", ["nocrop", "crop"])
def test_visibility_pipeline_vs_numba_fixture(tag):
    from deepviewagg_b200.core.multimodal import visibility as V
    g = load_golden("zbuffer_" + tag)
    W, H = [int(v) for v in g["size"]]
    ct, cb = [int(v) for v in g["crop"]]
    r_min, r_max = [float(v) for v in g["r"]]
    assert torch.equal(V.pose_to_rotation_matrix(g["img_opk"]), g["rotation"])
    idx, dist, xp, yp = V.camera_projection(g["xyz"].cuda(), g["img_xyz"], img_opk=g["img_opk"], img_size=(W, H),
                                            crop_top=ct, crop_bottom=cb, r_max=r_max, r_min=r_min)
    assert torch.equal(idx.cpu(), g["proj_idx"])                       # same kept set
    assert torch.equal(dist.cpu(), g["dist"])                          # float32 distances bit-exact
    # float64 pixel coordinates bit-identical to numba (FMA-chain sgemm + glibc atan2f / acosf, libm_f32.h)
    assert torch.equal(xp.cpu(), g["x_proj"]) and torch.equal(yp.cpu(), g["y_proj"])
    # integer stages on the reference's own projections: bit-exact
    xr, yr, dr = g["x_proj"].cuda(), g["y_proj"].cuda(), g["dist"].cuda()
    sp = V.splat_boxes(xr, yr, dr, None, (W, H), ct, cb, voxel=0.05, k_swell=1.0, d_swell=1000)
    assert torch.equal(sp.cpu(), g["splat"].int())
    for exact in (0, 1):
        i2, x2, y2 = V.visibility_from_splatting(xr, yr, dr, None, img_size=(W, H), crop_top=ct, crop_bottom=cb,
                                                 voxel=0.05, k_swell=1.0, d_swell=1000, exact=bool(exact))
        assert torch.equal(i2.cpu(), g[f"vis_idx_{exact}"])
        assert torch.equal(x2.cpu(), g[f"vis_x_{exact}"])
        assert torch.equal(y2.cpu(), g[f"vis_y_{exact}"])
    # whole pipeline from raw points (own projection -> boxes -> z-buffer): still the reference's winners
    for exact in (0, 1):
        i2, x2, y2 = V.visibility_from_splatting(xp, yp, dist, None, img_size=(W, H), crop_top=ct, crop_bottom=cb,
                                                 voxel=0.05, k_swell=1.0, d_swell=1000, exact=bool(exact))
        assert torch.equal(i2.cpu(), g[f"vis_idx_{exact}"]) and torch.equal(x2.cpu(), g[f"vis_x_{exact}"])


def test_pinhole_splat_vs_numba_fixture():
    from deepviewagg_b200.core.multimodal import visibility as V
    g = load_golden("splat_pinhole")
    W, H = [int(v) for v in g["size"]]
    intr = [[float(g["fx"]), 0, 0], [0, float(g["fy"]), 0]]
    sp = V.splat_boxes(g["x_proj"].cuda(), g["y_proj"].cuda(), g["dist"].cuda(), intr, (W, H), voxel=0.03,
                       k_swell=1.0, d_swell=1000, camera="scannet")
    assert torch.equal(sp.cpu(), g["splat"].int())


def test_zbuffer_large_random_vs_c_oracle():
    """1M points on a 2048x1024 equirectangular image (the BASELINE splat case), ties included."""
    from deepviewagg_b200.core.multimodal import visibility as V
    rng = np.random.default_rng(0)
    m, W, H = 1_000_000, 2048, 1024
    xp = rng.uniform(0, W, m)
    yp = rng.uniform(20, H - 20, m)
    dist = rng.uniform(0.6, 20, m).astype(np.float32)
    tie = np.arange(0, m - 7, 7)
    dist[tie] = dist[tie + 3]                                     # exact depth ties -> lowest index wins
    sp = VO.splat_boxes(xp, yp, dist, W, H, 8, 8, voxel=0.05)
    sg = V.splat_boxes(torch.from_numpy(xp).cuda(), torch.from_numpy(yp).cuda(), torch.from_numpy(dist).cuda(),
                       None, (W, H), 8, 8, voxel=0.05)
    assert np.array_equal(_np(sg), sp)
    for exact in (False, True):
        i_ref, x_ref, y_ref, _ = VO.zbuffer(sp, dist, xp, yp, W, H, 8, 8, exact=exact)
        i2, x2, y2 = V.visibility_from_splatting(torch.from_numpy(xp).cuda(), torch.from_numpy(yp).cuda(),
                                                 torch.from_numpy(dist).cuda(), None, img_size=(W, H), crop_top=8,
                                                 crop_bottom=8, voxel=0.05, exact=exact)
        assert np.array_equal(_np(i2), i_ref) and np.array_equal(_np(x2), x_ref) and np.array_equal(_np(y2), y_ref)


def test_csr_kernels_vs_oracle():
    from deepviewagg_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1)
    for n, ng in ((0, 5), (1, 1), (5000, 700), (100000, 100000), (10, 1000)):
        ids = np.sort(rng.integers(0, ng, n)).astype(np.int64)
        ref = VO.pointers_from_sorted_with_empties(ids, ng) if n else np.zeros(ng + 1, np.int64)
        d_ids = torch.from_numpy(ids).cuda()
        ptr = torch.full((ng + 1,), -7, dtype=torch.int64, device="cuda")
        _lib.check(lib.dva_csr_pointers_from_sorted(_lib.ptr(d_ids), _lib.ptr(ptr), n, ng, _lib.stream_ptr()), "csr")
        assert np.array_equal(_np(ptr), ref), (n, ng)
    g = load_golden("image_mapping")
    pointers, sel = g["pointers"], g["sel"]
    pn_ref, val_ref = VO.index_select_pointers(_np(pointers), _np(sel))
    assert np.array_equal(pn_ref, _np(g["sel_pointers"]))
    val = torch.empty(int(pn_ref[-1]), dtype=torch.int64, device="cuda")
    d_ptr, d_sel, d_pn = pointers.cuda(), sel.cuda(), torch.from_numpy(pn_ref).cuda()  # keep alive
    _lib.check(lib.dva_csr_select_values(_lib.ptr(d_ptr), _lib.ptr(d_sel), _lib.ptr(d_pn), _lib.ptr(val),
                                         sel.numel(), val.numel(), _lib.stream_ptr()), "select")
    assert np.array_equal(_np(val), val_ref)
    assert torch.equal(g["images"][val.cpu()], g["sel_images"])


@pytest.mark.parametrize("cam", ["scannet", "kitti360_perspective", "kitti360_fisheye"])
def test_pinhole_fisheye_cameras_vs_numba_fixture(cam):
    from deepviewagg_b200.core.multimodal import visibility as V
    g = load_golden("camera_" + cam)
    W, H = [int(v) for v in g["size"]]
    ct, cb = [int(v) for v in g["crop"]]
    fish = cam == "kitti360_fisheye"
    idx, dist, xp, yp = V.camera_projection(
        g["xyz"].cuda(), g["img_xyz"], img_intrinsic_pinhole=None if fish else g["pin"],
        img_intrinsic_fisheye=g["fish"] if fish else None, img_extrinsic=g["ext"], img_size=(W, H), crop_top=ct,
        crop_bottom=cb, r_max=float(g["r"][1]), r_min=float(g["r"][0]), camera=cam)
    assert torch.equal(idx.cpu(), g["proj_idx"]) and torch.equal(dist.cpu(), g["dist"])
    assert torch.equal(xp.cpu(), g["x_proj"]) and torch.equal(yp.cpu(), g["y_proj"])      # float64, bit-exact
    # image mask: drop the left half
    mask = torch.ones(W, H, dtype=torch.bool)
    mask[: W // 2] = False
    idx_m, _, xm, _ = V.camera_projection(
        g["xyz"].cuda(), g["img_xyz"], img_intrinsic_pinhole=None if fish else g["pin"],
        img_intrinsic_fisheye=g["fish"] if fish else None, img_extrinsic=g["ext"], img_mask=mask, img_size=(W, H),
        crop_top=ct, crop_bottom=cb, r_max=float(g["r"][1]), r_min=float(g["r"][0]), camera=cam)
    assert torch.equal(idx_m.cpu(), g["proj_idx"][g["x_proj"] >= W // 2]) and (xm >= W // 2).all()
    if fish:
        xr, yr, dr = g["x_proj"].cuda(), g["y_proj"].cuda(), g["dist"].cuda()
        xyz_kept = g["xyz"][g["proj_idx"]].cuda()
        sp = V.fisheye_splat_boxes(xr, yr, xyz_kept, g["ext"], g["fish"], (W, H), voxel=0.05)
        assert torch.equal(sp.cpu(), g["splat"].int())                       # every box, bit-exact
        for exact in (0, 1):
            i2, x2, y2 = V.visibility_from_splatting(xr, yr, dr, xyz_kept, img_extrinsic=g["ext"],
                                                     img_intrinsic_fisheye=g["fish"], img_size=(W, H), voxel=0.05,
                                                     exact=bool(exact), camera=cam)
            assert torch.equal(i2.cpu(), g[f"vis_idx_{exact}"])
            assert torch.equal(x2.cpu(), g[f"vis_x_{exact}"]) and torch.equal(y2.cpu(), g[f"vis_y_{exact}"])


@pytest.mark.parametrize("tag", ["equirect_exact", "equirect_splat", "scannet", "kitti360_fisheye"])
def test_splatting_visibility_dict_vs_reference(tag):
    """Z4 / Z5: the assembled SplattingVisibility.__call__ dict (visibility.py:1677-1776) against the
    executed reference (numba path): idx / x / y / depth bit-exact, postprocess_features (:1548-1582)
    within 1e-6 absolute (the four copied / affine columns bit-exact)."""
    import os
    from deepviewagg_b200.core.multimodal import visibility as V
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"visibility_model_{tag}.npz"))
    ctor = {k: (z["ctor/" + k].tolist() if z["ctor/" + k].ndim else z["ctor/" + k].item())
            for k in z["ctor_keys"].tolist()}
    ctor["img_size"] = tuple(ctor["img_size"])
    call = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("call/")}
    ref = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out/")}
    geo = torch.from_numpy(z["geo"]).cuda()
    model = V.SplattingVisibility(**ctor)
    out = model(torch.from_numpy(z["xyz"]).cuda(), torch.from_numpy(z["img_xyz"]), linearity=geo[:, 0],
                planarity=geo[:, 1], scattering=geo[:, 2], normals=torch.from_numpy(z["normals"]).cuda(), **call)
    for k in ("idx", "x", "y", "depth"):
        assert out[k].dtype == ref[k].dtype and torch.equal(out[k].cpu(), ref[k]), k
    f = out["features"].cpu()
    assert f.dtype == torch.float32 and f.shape == ref["features"].shape
    assert (f - ref["features"]).abs().max() <= 1e-6
    assert torch.equal(f[:, :4], ref["features"][:, :4])
    # the free function on the reference's own intermediate values
    sel = ref["idx"].cuda()
    xyz_to_img = torch.from_numpy(z["xyz"]).cuda()[sel] - torch.from_numpy(z["img_xyz"]).cuda()
    f2 = V.postprocess_features(xyz_to_img, None, out["depth"], geo[sel, 0], geo[sel, 1], geo[sel, 2],
                                torch.from_numpy(z["normals"]).cuda()[sel], **ctor)
    assert (f2.cpu() - ref["features"][:, :5]).abs().max() <= 1e-6


def test_map_images_equals_per_image_oracle():
    """MapImages (image.py transform :162-428) == per-image C-oracle visibility + numpy from_dense."""
    from deepviewagg_b200.core.multimodal.image import SameSettingImageData
    from deepviewagg_b200.core.multimodal.mapping import MapImages
    g = load_golden("zbuffer_nocrop")
    W, H = [int(v) for v in g["size"]]
    xyz = g["xyz"]
    cams = torch.stack([g["img_xyz"], g["img_xyz"] + torch.tensor([1.5, -0.7, 0.1]), torch.tensor([50., 50., 50.])])
    opk = torch.stack([g["img_opk"], g["img_opk"] * 0.5, g["img_opk"]])
    images = SameSettingImageData(pos=cams, opk=opk, ref_size=(W // 2, H // 2), proj_upscale=2, downscale=1)
    out = MapImages(voxel=0.05, exact=True, r_max=8, r_min=0.5)(xyz, images)
    assert out.num_views == 2                                  # the far-away third camera sees nothing
    m = out.mappings
    assert m.num_groups == xyz.shape[0] and m.pixels.dtype == torch.int16 and m.features.shape[1] == 2
    # oracle: same pipeline on the CPU
    pid, iid, pix = [], [], []
    for i in range(2):
        R = VO.pose_to_rotation_matrix(_np(opk[i]))
        dist, xp, yp, keep = VO.project_equirect(_np(xyz), _np(cams[i]), R, W, H, 0, 0, 0.5, 8.0)
        idx = np.where(keep)[0]
        sp = VO.splat_boxes(xp[idx], yp[idx], dist[idx], W, H, voxel=0.05)
        i2, x2, y2, _ = VO.zbuffer(sp, dist[idx], xp[idx], yp[idx], W, H, exact=True)
        p, x, y = idx[i2], x2 // 2, y2 // 2
        u = VO.lexargunique(p, x, y)
        pid.append(p[u]); iid.append(np.full(len(u), i)); pix.append(np.stack([x[u], y[u]], 1))
    ref = VO.image_mapping_from_dense(np.concatenate(pid), np.concatenate(iid), np.concatenate(pix), None,
                                      xyz.shape[0])
    assert np.array_equal(_np(m.pointers), ref["pointers"]) and np.array_equal(_np(m.images), ref["images"])
    assert np.array_equal(_np(m.atomic_csr_indexing), ref["atomic_pointers"])
    assert np.array_equal(_np(m.pixels).astype(np.int64), ref["pixels"])


def _mapping_equal(a, b):
    """two ImageMappings (any device): integers bit-exact, pixels in the same order, features to 1e-6"""
    assert torch.equal(a.pointers.cpu(), b.pointers.cpu())
    assert torch.equal(a.images.cpu(), b.images.cpu())
    assert torch.equal(a.values[1].pointers.cpu(), b.values[1].pointers.cpu())
    assert a.pixels.dtype == b.pixels.dtype and torch.equal(a.pixels.cpu(), b.pixels.cpu())
    if a.has_features or b.has_features:
        assert torch.allclose(a.features.cpu(), b.features.cpu(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n_points,n_items,n_img,pix_dtype", [(5000, 60000, 7, torch.int16), (300, 40000, 3, torch.int32),
                                                            (100000, 300000, 40, torch.int16), (1, 1, 1, torch.int64)])
def test_native_mapping_build_equals_host_path(n_points, n_items, n_img, pix_dtype):
    """dva_mapping_build (bucket by point + warp rank sort; csrc/mapping_build.cu) == the torch-sort host path
    (itself pinned on the executed reference, tests/test_containers.py) for from_dense and for
    select_points('merge'): unseen points, multi-pixel views, buckets longer than a warp (300 points x
    40 000 items -> ~130 items per point), duplicated (point, image, pixel) triples."""
    from deepviewagg_b200.core.multimodal.image import ImageMapping
    gen = torch.Generator().manual_seed(n_items + n_points)
    pid = torch.randint(0, max(n_points - 3, 1), (n_items,), generator=gen)            # the last points stay unseen
    iid = torch.randint(0, n_img, (n_items,), generator=gen)
    pix = torch.randint(0, 50, (n_items, 2), generator=gen).to(pix_dtype)
    feat = torch.rand(n_items, 8, generator=gen)
    ref = ImageMapping.from_dense(pid, iid, pix, feat, num_points=n_points)             # CPU tensors: torch sorts
    got = ImageMapping.from_dense(pid.cuda(), iid.cuda(), pix.cuda(), feat.cuda(), num_points=n_points)
    _mapping_equal(got, ref)
    nof = ImageMapping.from_dense(pid.cuda(), iid.cuda(), pix.cuda(), None, num_points=n_points)
    assert not nof.has_features and torch.equal(nof.images.cpu(), ref.images)
    # merge: agglomerate points 4 -> 1 (every output voxel present), duplicates removed, features averaged per view
    n_out = max(n_points // 4, 1)
    idx = torch.randint(0, n_out, (n_points,), generator=gen)
    idx[:n_out] = torch.arange(n_out)
    ref_m = ref.select_points(idx, mode="merge")
    got_m = got.select_points(idx.cuda(), mode="merge")
    assert torch.equal(got_m.pointers.cpu(), ref_m.pointers) and torch.equal(got_m.images.cpu(), ref_m.images)
    ap = ref_m.values[1].pointers
    assert torch.equal(got_m.values[1].pointers.cpu(), ap)
    from test_containers import canon_pixels
    assert torch.equal(canon_pixels(got_m.pixels.cpu(), ap), canon_pixels(ref_m.pixels, ap))
    assert torch.allclose(got_m.features.cpu(), ref_m.features, rtol=1e-5, atol=1e-6)
    with pytest.raises(IndexError):
        ImageMapping.from_dense(pid.cuda() + n_points, iid.cuda(), pix.cuda(), None, num_points=n_points)
    # no item at all: every point unseen, empty level-2 CSR
    e = ImageMapping.from_dense(pid[:0].cuda(), iid[:0].cuda(), pix[:0].cuda(), None, num_points=n_points)
    assert e.num_groups == n_points and int(e.pointers.abs().sum()) == 0 and e.images.numel() == 0


def test_view_cat_sorting_closed_form_equals_stable_argsort():
    from deepviewagg_b200.core.multimodal.image import ImageData, ImageMapping, SameSettingImageData
    gen = torch.Generator().manual_seed(12)
    N = 20000
    ims = []
    for s, (n_img, n_items) in enumerate(((4, 90000), (2, 30000), (3, 1), (5, 150000))):
        pid = torch.randint(0, N, (n_items,), generator=gen)
        iid = torch.randint(0, n_img, (n_items,), generator=gen)
        pix = torch.randint(0, 32, (n_items, 2), generator=gen).short()
        im = SameSettingImageData(pos=torch.zeros(n_img, 3), opk=torch.zeros(n_img, 3), ref_size=(32 + s, 32),
                                  proj_upscale=1, downscale=1)
        im.mappings = ImageMapping.from_dense(pid, iid, pix, torch.rand(n_items, 8, generator=gen), num_points=N)
        ims.append(im)
    cpu = ImageData(ims)
    gpu = cpu.to("cuda")
    assert torch.equal(gpu.view_cat_sorting.cpu(), cpu.view_cat_sorting)            # stable order of equal points
    assert torch.equal(gpu.view_cat_csr_indexing.cpu(), cpu.view_cat_csr_indexing)
    srt, csr = gpu._view_cat_native()
    assert torch.equal(csr.cpu(), cpu.view_cat_csr_indexing)
