"""ctypes wrapper of oracle/visibility_oracle.c plus numpy restatements of the integer CSR /
lexicographic helpers (oracle; test infrastructure -- see oracle/__init__.py).

  project_equirect / splat_boxes / zbuffer   <- visibility.py (numba CPU variants), see the .c file
  pose_to_rotation_matrix                    <- visibility.py:57-90
  lexargsort / lexargunique / lexunique      <- utils/multimodal.py:36-94, 289-323 (CPU/numpy path;
                                                argsort made stable: the reference's np.argsort is
                                                unstable, SURVEY.md D.15 -- ties are canonicalised)
  csr_pointers / insert_empty_groups / index_select_pointers <- csr.py:158-264
  image_mapping_from_dense                   <- image.py:1728-1795
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = ctypes.CDLL(_LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pose_to_rotation_matrix(opk):
    """visibility.py:57-90, float32 like numba (cos/sin of float32 scalars, float32 products)."""
    opk = np.asarray(opk, dtype=np.float32)
    co, so, cp, sp, ck, sk = (np.cos(opk[0]), np.sin(opk[0]), np.cos(opk[1]), np.sin(opk[1]),
                              np.cos(opk[2]), np.sin(opk[2]))
    M_o = np.array([[1, 0, 0], [0, co, -so], [0, so, co]], dtype=np.float32)
    M_p = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float32)
    M_k = np.array([[ck, -sk, 0], [sk, ck, 0], [0, 0, 1]], dtype=np.float32)
    return np.dot(M_o, np.dot(M_p, M_k)).astype(np.float32)


def project_equirect(xyz, img_xyz, rotation, W, H, crop_top=0, crop_bottom=0, r_min=0.5, r_max=30.0):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    n = xyz.shape[0]
    dist = np.empty(n, np.float32)
    xp, yp = np.empty(n, np.float64), np.empty(n, np.float64)
    keep = np.empty(n, np.uint8)
    R = np.ascontiguousarray(rotation, dtype=np.float32).reshape(-1)
    c = np.ascontiguousarray(img_xyz, dtype=np.float32)
    _load().oracle_project_equirect(_p(xyz), _p(c), _p(R), ctypes.c_int64(n), W, H, crop_top, crop_bottom,
                                    ctypes.c_float(r_min), ctypes.c_float(r_max), _p(dist), _p(xp), _p(yp),
                                    _p(keep))
    return dist, xp, yp, keep.astype(bool)


def camera_transform(camera, img_extrinsic):
    """(A, t0, t1) with p = A (xyz - t0) + t1, float32 (visibility.py:231-244, 304-310)."""
    E = np.ascontiguousarray(np.asarray(img_extrinsic, dtype=np.float32))
    if camera == "scannet":
        # numba's np.linalg.inv = LAPACK sgetrf + sgetri (numpy's inv uses sgesv: last-bit differences)
        from scipy.linalg import lapack
        lu, piv, _ = lapack.sgetrf(E)
        c2w = np.ascontiguousarray(lapack.sgetri(lu, piv)[0], dtype=np.float32)
        return c2w[:3, :3].copy(), np.zeros(3, np.float32), c2w[:3, 3].copy()
    return E[:3, :3].T.copy(), E[:3, 3].copy(), np.zeros(3, np.float32)


def project_camera(xyz, img_xyz, camera, img_extrinsic, intrinsic, W, H, crop_top=0, crop_bottom=0,
                   r_min=0.5, r_max=30.0):
    """Pinhole ('scannet', 'kitti360_perspective': intrinsic = 4x4 matrix) or fisheye
    ('kitti360_fisheye': intrinsic = [xi,k1,k2,gamma1,gamma2,u0,v0]) projection."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    n = xyz.shape[0]
    A, t0, t1 = [np.ascontiguousarray(v, dtype=np.float32) for v in camera_transform(camera, img_extrinsic)]
    intr = np.zeros(8, np.float32)
    if camera == "kitti360_fisheye":
        intr[:7] = np.asarray(intrinsic, dtype=np.float32)
        cam = 3
    else:
        K = np.asarray(intrinsic, dtype=np.float32)
        intr[:4] = [K[0, 0], K[1, 1], K[0, 2], K[1, 2]]
        cam = 1
    dist = np.empty(n, np.float32)
    xp, yp = np.empty(n, np.float64), np.empty(n, np.float64)
    keep = np.empty(n, np.uint8)
    c = np.ascontiguousarray(img_xyz, dtype=np.float32)
    _load().oracle_project_camera(_p(xyz), _p(c), _p(A.reshape(-1)), _p(t0), _p(t1), _p(intr), cam,
                                  ctypes.c_int64(n), W, H, crop_top, crop_bottom, ctypes.c_float(r_min),
                                  ctypes.c_float(r_max), _p(dist), _p(xp), _p(yp), _p(keep))
    return dist, xp, yp, keep.astype(bool)


def fisheye_splat(x_proj, y_proj, xyz, img_extrinsic, fish, W, H, crop_top=0, crop_bottom=0, voxel=0.02,
                  k_swell=1.0, d_swell=1000.0):
    """fisheye_splat_cpu, visibility.py:876-953 (width from the projection of the voxel top;
    `dist = norm_cpu(xyz)` of the absolute coordinates, :900)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    d = np.sqrt((xyz ** 2).sum(axis=1)).astype(np.float32)
    swell = 1 + k_swell * np.exp(-d.astype(np.float64) / np.log(d_swell))
    top = xyz.copy()
    top[:, 2] += (swell * voxel / 2).astype(np.float32)
    _, xt, yt, _ = project_camera(top, np.zeros(3, np.float32), "kitti360_fisheye", img_extrinsic, fish,
                                  1 << 20, 1 << 20, 0, 0, 0.0, 1e30)
    w = 2 * np.sqrt((np.asarray(x_proj, np.float64) - xt) ** 2 + (np.asarray(y_proj, np.float64) - yt) ** 2)
    xa = np.rint(x_proj - w / 2).astype(np.float32).astype(np.int32)
    xb = np.rint(x_proj + w / 2 + 1).astype(np.float32).astype(np.int32)
    ya = np.rint(y_proj - w / 2).astype(np.float32).astype(np.int32)
    yb = np.rint(y_proj + w / 2 + 1).astype(np.float32).astype(np.int32)
    y_min, y_max = crop_top, H - crop_bottom
    return np.stack([np.clip(xa, 0, W - 1), np.clip(xb, 1, W), np.clip(ya, y_min, y_max - 1),
                     np.clip(yb, y_min + 1, y_max)], axis=1).astype(np.int32)


def splat_boxes(x_proj, y_proj, dist, W, H, crop_top=0, crop_bottom=0, voxel=0.02, k_swell=1.0,
                d_swell=1000.0, camera="equirectangular", fx=0.0, fy=0.0):
    xp = np.ascontiguousarray(x_proj, np.float64)
    yp = np.ascontiguousarray(y_proj, np.float64)
    d = np.ascontiguousarray(dist, np.float32)
    m = xp.shape[0]
    out = np.empty((m, 4), np.int32)
    D = ctypes.c_double
    if camera == "equirectangular":
        _load().oracle_splat_equirect(_p(xp), _p(yp), _p(d), ctypes.c_int64(m), W, H, crop_top, crop_bottom,
                                      D(voxel), D(k_swell), D(d_swell), _p(out))
    else:
        _load().oracle_splat_pinhole(_p(xp), _p(yp), _p(d), ctypes.c_int64(m), W, H, crop_top, crop_bottom,
                                     D(voxel), D(k_swell), D(d_swell), D(fx), D(fy), _p(out))
    return out


def zbuffer(splat, dist, x_proj, y_proj, W, H, crop_top=0, crop_bottom=0, exact=False):
    """-> (indices, x_pix, y_pix) in the reference's np.where order (row-major over [x, y])."""
    sp = np.ascontiguousarray(splat, np.int32)
    d = np.ascontiguousarray(dist, np.float32)
    xp = np.ascontiguousarray(x_proj, np.float64)
    yp = np.ascontiguousarray(y_proj, np.float64)
    m = sp.shape[0]
    Hc = H - crop_top - crop_bottom
    idx_map = np.empty((W, Hc), np.int64)
    _load().oracle_zbuffer(_p(sp), _p(d), _p(xp), _p(yp), ctypes.c_int64(m), W, H, crop_top, crop_bottom,
                           int(bool(exact)), _p(idx_map))
    x_pix, y_pix = np.where(idx_map != -1)
    return idx_map[x_pix, y_pix], x_pix, y_pix + crop_top, idx_map


# ---- lexicographic helpers (utils/multimodal.py) ---------------------------------------------------
def postprocess_features(xyz_to_img, y_proj, dist, linearity, planarity, scattering, normals, img_size,
                         r_max, r_min):
    """visibility.py:1548-1582 (+ normalize_dist_cuda :1503-1518, orientation_cuda :1521-1545): the
    [k, F] float32 viewing-condition features.  torch CPU float32 arithmetic restated in numpy."""
    f32 = np.float32
    feats = []
    if dist is not None:
        d = np.asarray(dist, f32)
        feats.append(((d - f32(r_min)) / f32(r_max + 1e-4)).astype(f32))
    for f in (linearity, planarity, scattering):
        if f is not None:
            feats.append(np.asarray(f, f32))
    if xyz_to_img is not None and dist is not None and normals is not None:
        u = (np.asarray(xyz_to_img, f32) / (np.asarray(dist, f32) + f32(1e-4)).reshape(-1, 1)).astype(f32)
        p = u * np.asarray(normals, f32)
        feats.append(np.abs((p[:, 0] + p[:, 1]) + p[:, 2]).astype(f32))
    if y_proj is not None:
        feats.append((np.asarray(y_proj, np.float64) / img_size[1]).astype(f32))
    return np.stack(feats).T


def splatting_visibility(xyz, img_xyz, linearity=None, planarity=None, scattering=None, normals=None,
                         img_opk=None, img_extrinsic=None, img_intrinsic_pinhole=None,
                         img_intrinsic_fisheye=None, img_size=(1024, 512), crop_top=0, crop_bottom=0, r_max=30,
                         r_min=0.5, camera="s3dis_equirectangular", voxel=0.1, k_swell=1.0, d_swell=1000,
                         exact=False):
    """SplattingVisibility.__call__ (visibility.py:1677-1776): projection -> splat boxes -> z-buffer ->
    features, assembled exactly like VisibilityModel.__call__ :1694-1757."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    img_xyz = np.asarray(img_xyz, np.float32)
    W, H = int(img_size[0]), int(img_size[1])
    if camera == "s3dis_equirectangular":
        R = pose_to_rotation_matrix(np.asarray(img_opk, np.float32))
        dist, xp, yp, keep = project_equirect(xyz, img_xyz, R, W, H, crop_top, crop_bottom, r_min, r_max)
    else:
        intr = img_intrinsic_fisheye if camera == "kitti360_fisheye" else img_intrinsic_pinhole
        dist, xp, yp, keep = project_camera(xyz, img_xyz, camera, img_extrinsic, np.asarray(intr), W, H, crop_top,
                                            crop_bottom, r_min, r_max)
    idx_1 = np.where(keep)[0]
    if idx_1.size == 0:
        e = np.zeros(0, np.int64)
        return dict(idx=e, x=e.copy(), y=e.copy(), depth=np.zeros(0, np.float32), features=np.zeros(0, np.float32))
    dist, xp, yp = dist[idx_1], xp[idx_1], yp[idx_1]
    if camera == "s3dis_equirectangular":
        sp = splat_boxes(xp, yp, dist, W, H, crop_top, crop_bottom, voxel, k_swell, d_swell)
    elif camera == "kitti360_fisheye":
        sp = fisheye_splat(xp, yp, xyz[idx_1], img_extrinsic, np.asarray(img_intrinsic_fisheye), W, H, crop_top,
                           crop_bottom, voxel, k_swell, d_swell)
    else:
        K = np.asarray(img_intrinsic_pinhole, np.float32)
        sp = splat_boxes(xp, yp, dist, W, H, crop_top, crop_bottom, voxel, k_swell, d_swell, camera="pinhole",
                         fx=float(K[0, 0]), fy=float(K[1, 1]))
    idx_2, x_pix, y_pix, _ = zbuffer(sp, dist, xp, yp, W, H, crop_top, crop_bottom, exact=exact)
    idx = idx_1[idx_2]
    pick = lambda a: None if a is None else np.asarray(a)[idx]  # noqa: E731
    feats = postprocess_features(xyz[idx] - img_xyz, yp[idx_2], dist[idx_2], pick(linearity), pick(planarity),
                                 pick(scattering), pick(normals), (W, H), r_max, r_min)
    return dict(idx=idx, x=x_pix, y=y_pix, depth=dist[idx_2], features=feats)


def _composite(*arrays):
    """CompositeNDArray (utils/multimodal.py:175-250): key = sum a_i * prod_{j>i} (max_j + 1)."""
    arrays = [np.asarray(a).astype(np.int64) for a in arrays]
    if arrays[0].shape[0] == 0:
        return np.zeros(0, np.int64), [1] * len(arrays)
    maxs = [int(np.abs(a).max()) + 1 for a in arrays]
    bases = [int(np.prod(maxs[i + 1:])) for i in range(len(arrays) - 1)] + [1]
    return sum(a * b for a, b in zip(arrays, bases)), bases


def lexargsort(*arrays):
    key, _ = _composite(*arrays)
    return np.argsort(key, kind="stable")


def lexargunique(*arrays):
    key, _ = _composite(*arrays)
    return np.unique(key, return_index=True)[1]


def lexunique(*arrays):
    key, bases = _composite(*arrays)
    u = np.unique(key)
    out = []
    for b in bases:
        out.append(u // b)
        u = u % b
    return out


# ---- CSR containers (csr.py) ---------------------------------------------------------------------------
def csr_pointers(sorted_ids):
    """_sorted_indices_to_pointers, csr.py:158-172."""
    ids = np.asarray(sorted_ids)
    return np.concatenate([[0], np.where(ids[1:] > ids[:-1])[0] + 1, [ids.shape[0]]]).astype(np.int64)


def insert_empty_groups(pointers, group_indices, num_groups=None):
    """csr.py:197-229: pointers.repeat_interleave(ends - starts)."""
    gi = np.asarray(group_indices).astype(np.int64)
    ng = int(gi.max()) + 1 if num_groups is None else max(int(gi.max()) + 1, int(num_groups))
    starts = np.concatenate([[-1], gi])
    ends = np.concatenate([gi, [ng]])
    return np.repeat(np.asarray(pointers), ends - starts)


def pointers_from_sorted_with_empties(sorted_ids, num_groups):
    ids = np.asarray(sorted_ids)
    p = csr_pointers(ids)
    return insert_empty_groups(p, ids[p[1:] - 1], num_groups)


def index_select_pointers(pointers, indices):
    """csr.py:235-264 -> (pointers_new, val_idx)."""
    pointers, indices = np.asarray(pointers), np.asarray(indices)
    sizes = pointers[indices + 1] - pointers[indices]
    pn = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    val = np.arange(pn[-1]) - np.repeat(pn[:-1], sizes) + np.repeat(pointers[indices], sizes)
    return pn, val.astype(np.int64)


def image_mapping_from_dense(point_ids, image_ids, pixels, features, num_points):
    """ImageMapping.from_dense, image.py:1728-1795 (stable sort; see module docstring)."""
    order = lexargsort(point_ids, image_ids)
    pid, iid, pix = np.asarray(point_ids)[order], np.asarray(image_ids)[order], np.asarray(pixels)[order]
    feat = None if features is None else np.asarray(features)[order]
    key, _ = _composite(pid, iid)
    atomic_ptr = csr_pointers(key)
    last = atomic_ptr[1:] - 1
    iid_v, pid_v = iid[last], pid[last]
    if feat is not None:
        counts = np.diff(atomic_ptr)
        sums = np.add.reduceat(feat.astype(np.float32), atomic_ptr[:-1], axis=0)
        feat = (sums / np.maximum(counts, 1)[:, None]).astype(np.float32)
    ptr = csr_pointers(pid_v)
    n = max(int(num_points), int(pid_v.max()) + 1)
    ptr = insert_empty_groups(ptr, pid_v[ptr[1:] - 1], n)
    return dict(pointers=ptr, images=iid_v, atomic_pointers=atomic_ptr, pixels=pix, features=feat)
