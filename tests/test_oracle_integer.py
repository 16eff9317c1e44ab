"""Pin the integer side of the oracle (oracle/visibility_oracle.{c,py}) against fixtures produced by
the EXECUTED reference: numba CPU visibility (visibility.py), ImageMapping.from_dense / indexing
(image.py, csr.py) and the lex helpers (utils/multimodal.py).  Bit-exact. CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import visibility_oracle as VO


def _np(t):
    return t.numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def test_lex_kat():
    g = load_golden("kat_lex")
    a, b = _np(g["a"]), _np(g["b"])
    s = VO.lexargsort(a, b)
    # reference argsort is unstable: compare the sorted keys, and the reference's own order
    assert (a[s].tolist(), b[s].tolist()) == ([0, 0, 1, 2, 2], [5, 5, 3, 0, 1])
    rs = _np(g["argsort"])
    assert (a[rs].tolist(), b[rs].tolist()) == ([0, 0, 1, 2, 2], [5, 5, 3, 0, 1])
    assert VO.lexargunique(a, b).tolist() == _np(g["argunique"]).tolist() == [1, 3, 2, 0]
    u = VO.lexunique(a, b)
    assert u[0].tolist() == _np(g["unique_a"]).tolist() and u[1].tolist() == _np(g["unique_b"]).tolist()


def test_image_mapping_from_dense_bit_exact():
    g = load_golden("image_mapping")
    m = VO.image_mapping_from_dense(_np(g["point_ids"]), _np(g["image_ids"]), _np(g["pixels"]),
                                    _np(g["features"]), int(g["num_points"]))
    assert np.array_equal(m["pointers"], _np(g["pointers"]))
    assert np.array_equal(m["images"], _np(g["images"]))
    assert np.array_equal(m["atomic_pointers"], _np(g["atomic_pointers"]))
    # pixels of one (point,image) view may come in any order (unstable sort): canonicalise
    ap = m["atomic_pointers"]

    def canon(pix):
        out = pix.copy()
        for i in range(len(ap) - 1):
            seg = out[ap[i]:ap[i + 1]]
            out[ap[i]:ap[i + 1]] = seg[np.lexsort((seg[:, 1], seg[:, 0]))]
        return out
    assert np.array_equal(canon(m["pixels"]), canon(_np(g["out_pixels"])))
    assert np.allclose(m["features"], _np(g["out_features"]), rtol=1e-6, atol=1e-7)
    # CSR group selection (csr.py:235-264)
    pn, val = VO.index_select_pointers(m["pointers"], _np(g["sel"]))
    assert np.array_equal(pn, _np(g["sel_pointers"]))
    assert np.array_equal(m["images"][val], _np(g["sel_images"]))


@pytest.mark.parametrize("tag", ["nocrop", "crop"])
def test_projection_splat_zbuffer_vs_numba(tag):
    g = load_golden("zbuffer_" + tag)
    W, H = [int(v) for v in g["size"]]
    ct, cb = [int(v) for v in g["crop"]]
    r_min, r_max = [float(v) for v in g["r"]]
    R = VO.pose_to_rotation_matrix(_np(g["img_opk"]))
    assert np.array_equal(R, _np(g["rotation"]))
    dist, xp, yp, keep = VO.project_equirect(_np(g["xyz"]), _np(g["img_xyz"]), R, W, H, ct, cb, r_min, r_max)
    idx = np.where(keep)[0]
    ref_idx = _np(g["proj_idx"])
    # float projection: the set of kept points and their integer pixels must agree; report the rate
    same_set = np.array_equal(idx, ref_idx)
    assert same_set, f"kept sets differ: {len(idx)} vs {len(ref_idx)}"
    assert np.array_equal(dist[idx], _np(g["dist"]))
    # float64 pixel coordinates bit-identical to numba (sgemm FMA chain + libm atan2f / acosf)
    assert np.array_equal(xp[idx], _np(g["x_proj"])) and np.array_equal(yp[idx], _np(g["y_proj"]))
    # from here on integers only, bit-exact given the reference's projections
    xr, yr, dr = _np(g["x_proj"]), _np(g["y_proj"]), _np(g["dist"])
    sp = VO.splat_boxes(xr, yr, dr, W, H, ct, cb, voxel=0.05, k_swell=1.0, d_swell=1000)
    assert np.array_equal(sp, _np(g["splat"]))
    for exact in (0, 1):
        i2, x2, y2, _ = VO.zbuffer(sp, dr, xr, yr, W, H, ct, cb, exact=bool(exact))
        assert np.array_equal(i2, _np(g[f"vis_idx_{exact}"]))
        assert np.array_equal(x2, _np(g[f"vis_x_{exact}"]))
        assert np.array_equal(y2, _np(g[f"vis_y_{exact}"]))


def test_pinhole_splat_vs_numba():
    g = load_golden("splat_pinhole")
    W, H = [int(v) for v in g["size"]]
    sp = VO.splat_boxes(_np(g["x_proj"]), _np(g["y_proj"]), _np(g["dist"]), W, H, voxel=0.03, k_swell=1.0,
                        d_swell=1000, camera="pinhole", fx=float(g["fx"]), fy=float(g["fy"]))
    assert np.array_equal(sp, _np(g["splat"]))


@pytest.mark.parametrize("cam", ["scannet", "kitti360_perspective", "kitti360_fisheye"])
def test_pinhole_fisheye_projection_vs_numba(cam):
    g = load_golden("camera_" + cam)
    W, H = [int(v) for v in g["size"]]
    ct, cb = [int(v) for v in g["crop"]]
    intr = _np(g["fish"]) if cam == "kitti360_fisheye" else _np(g["pin"])
    d, xp, yp, keep = VO.project_camera(_np(g["xyz"]), _np(g["img_xyz"]), cam, _np(g["ext"]), intr, W, H, ct, cb,
                                        float(g["r"][0]), float(g["r"][1]))
    idx = np.where(keep)[0]
    assert np.array_equal(idx, _np(g["proj_idx"]))                    # same kept set
    assert np.array_equal(d[idx], _np(g["dist"]))                     # float32 distances bit-exact
    assert np.array_equal(xp[idx], _np(g["x_proj"])) and np.array_equal(yp[idx], _np(g["y_proj"]))  # f64, bit-exact
    if cam == "kitti360_fisheye":
        xr, yr, dr = _np(g["x_proj"]), _np(g["y_proj"]), _np(g["dist"])
        sp = VO.fisheye_splat(xr, yr, _np(g["xyz"])[idx], _np(g["ext"]), intr, W, H, voxel=0.05)
        ref = _np(g["splat"])
        assert np.array_equal(sp, ref)                                 # width from a second float32 projection
        for exact in (0, 1):                                          # z-buffer on the reference's own boxes
            i2, x2, y2, _ = VO.zbuffer(ref, dr, xr, yr, W, H, exact=bool(exact))
            assert np.array_equal(i2, _np(g[f"vis_idx_{exact}"]))
            assert np.array_equal(x2, _np(g[f"vis_x_{exact}"])) and np.array_equal(y2, _np(g[f"vis_y_{exact}"]))


def _load_vis_model(tag):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"visibility_model_{tag}.npz"))
    ctor = {k: (z["ctor/" + k].tolist() if z["ctor/" + k].ndim else z["ctor/" + k].item()) for k in z["ctor_keys"].tolist()}
    ctor["img_size"] = tuple(ctor["img_size"])
    call = {k[5:]: z[k] for k in z.files if k.startswith("call/")}
    out = {k[4:]: z[k] for k in z.files if k.startswith("out/")}
    return z, ctor, call, out


@pytest.mark.parametrize("tag", ["equirect_exact", "equirect_splat", "scannet", "kitti360_fisheye"])
def test_splatting_visibility_dict_vs_reference(tag):
    """Z4 / Z5: the whole SplattingVisibility.__call__ dict of the executed reference (numba path):
    idx / x / y / depth bit-exact, the six postprocess_features columns to float32 rounding."""
    z, ctor, call, ref = _load_vis_model(tag)
    geo = z["geo"]
    out = VO.splatting_visibility(z["xyz"], z["img_xyz"], linearity=geo[:, 0], planarity=geo[:, 1],
                                  scattering=geo[:, 2], normals=z["normals"], **ctor, **call)
    for k in ("idx", "x", "y"):
        assert np.array_equal(out[k], ref[k]), k
    assert np.array_equal(out["depth"], ref["depth"])
    assert out["features"].shape == ref["features"].shape and out["features"].dtype == np.float32
    assert np.abs(out["features"] - ref["features"]).max() <= 1e-6
    assert np.array_equal(out["features"][:, :4], ref["features"][:, :4])      # depth + geometric columns: exact
