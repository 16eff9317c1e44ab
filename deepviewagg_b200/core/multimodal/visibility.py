"""Point -> pixel visibility on the GPU, with the reference's function / class names
(torch_points3d/core/multimodal/visibility.py).  Results follow the reference's CPU (numba)
variants, which are the authoritative ones (its README.md:122-123 warns against its own GPU
mapping path): integer outputs -- splat boxes, z-buffer winners, pixel coordinates -- are
bit-identical to the numba loops given the same projections.

  camera_projection          <- visibility.py:478-538, 592-623   (equirectangular, pinhole, fisheye)
  visibility_from_splatting  <- visibility.py:1073-1195, 1288-1322
  postprocess_features       <- visibility.py:1548-1582
  VisibilityModel, SplattingVisibility <- visibility.py:1677-1776

Kernels: csrc/zbuffer.cu through the C ABI (dva_project_equirectangular, dva_project_camera,
dva_splat_boxes, dva_splat_boxes_from_width, dva_zbuffer_splat).  Compaction of the kept set / winner map is index plumbing (torch.nonzero).
Depth-map and Biasutti visibility models are out of scope (not used by any shipped config).
"""
import ctypes

import numpy as np
import torch

from ... import _lib
from ..._lib import check, ptr, require_cuda, stream_ptr

_PINHOLE_CAMERAS = ("scannet", "kitti360_perspective")


def pose_to_rotation_matrix(opk):
    """Rotation matrix of an (omega, phi, kappa) pose: M_o . (M_p . M_k), float32 -- the
    arithmetic of pose_to_rotation_matrix_cpu (visibility.py:57-90), done on the host."""
    opk = np.asarray(opk.detach().cpu().numpy() if isinstance(opk, torch.Tensor) else opk, dtype=np.float32)
    co, so = np.cos(opk[0]), np.sin(opk[0])
    cp, sp = np.cos(opk[1]), np.sin(opk[1])
    ck, sk = np.cos(opk[2]), np.sin(opk[2])
    m_o = np.array([[1.0, 0.0, 0.0], [0.0, co, -so], [0.0, so, co]], dtype=np.float32)
    m_p = np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]], dtype=np.float32)
    m_k = np.array([[ck, -sk, 0.0], [sk, ck, 0.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    return torch.from_numpy(np.dot(m_o, np.dot(m_p, m_k)).astype(np.float32))


def _inv4_f32(E):
    """float32 inverse of the 4x4 extrinsic exactly as numba's np.linalg.inv computes it
    (visibility.py:233): LAPACK sgetrf + sgetri (numpy's own inv solves against the identity with
    sgesv instead and differs in the last bit, which moves ~40 % of the float pixel coordinates
    by one ulp).  Host-side camera set-up, once per image."""
    try:
        from scipy.linalg import lapack
    except ImportError as e:  # pragma: no cover - scipy ships with the image (numba needs it too)
        raise ImportError("the 'scannet' camera needs scipy's LAPACK bindings for a bit-exact "
                          "camera-to-world matrix") from e
    lu, piv, info = lapack.sgetrf(E)
    if info != 0:
        raise np.linalg.LinAlgError("singular extrinsic matrix")
    inv, info = lapack.sgetri(lu, piv)
    if info != 0:
        raise np.linalg.LinAlgError("singular extrinsic matrix")
    return np.ascontiguousarray(inv, dtype=np.float32)


def _camera_transform(camera, img_extrinsic):
    """(A, t0, t1) float32 with p = A (xyz - t0) + t1 (visibility.py:231-244, 304-310)."""
    E = np.ascontiguousarray(np.asarray(
        img_extrinsic.detach().cpu().numpy() if isinstance(img_extrinsic, torch.Tensor) else img_extrinsic,
        dtype=np.float32))
    if camera == 'scannet':
        c2w = _inv4_f32(E)
        return c2w[:3, :3].copy(), np.zeros(3, np.float32), c2w[:3, 3].copy()
    return E[:3, :3].T.copy(), E[:3, 3].copy(), np.zeros(3, np.float32)


def _host_f32(t, n):
    t = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    return np.asarray(t, dtype=np.float32).reshape(-1)[:n]


def camera_projection(xyz, img_xyz, img_opk=None, img_intrinsic_pinhole=None, img_intrinsic_fisheye=None,
                      img_extrinsic=None, img_mask=None, img_size=(1024, 512), crop_top=0, crop_bottom=0,
                      r_max=30, r_min=0.5, camera='s3dis_equirectangular', **kwargs):
    """-> (indices int64[m], dist f32[m], x_proj f64[m], y_proj f64[m]) of the points within
    (r_min, r_max) of the camera that project inside the (cropped) image and its mask
    (visibility.py:478-538).  Cameras: s3dis_equirectangular, scannet, kitti360_perspective,
    kitti360_fisheye."""
    require_cuda(xyz)
    lib = _lib.load()
    dev = xyz.device
    xyz = xyz.float().contiguous()
    n = xyz.shape[0]
    W, H = int(img_size[0]), int(img_size[1])
    dist = torch.empty(n, dtype=torch.float32, device=dev)
    x_proj = torch.empty(n, dtype=torch.float64, device=dev)
    y_proj = torch.empty(n, dtype=torch.float64, device=dev)
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    cam_xyz = _host_f32(img_xyz, 3)
    with torch.cuda.device(dev):
        if camera == 's3dis_equirectangular':
            rot = pose_to_rotation_matrix(img_opk if img_opk is not None else np.zeros(3, np.float32))
            pose = torch.from_numpy(np.concatenate([cam_xyz, rot.numpy().reshape(-1)])).to(dev)
            check(lib.dva_project_equirectangular(ptr(xyz), ptr(pose), ptr(dist), ptr(x_proj), ptr(y_proj),
                                                  ptr(keep), n, W, H, int(crop_top), int(crop_bottom),
                                                  float(r_min), float(r_max), stream_ptr()),
                  "dva_project_equirectangular")
        elif camera in _PINHOLE_CAMERAS or camera == 'kitti360_fisheye':
            A, t0, t1 = _camera_transform(camera, img_extrinsic)
            intr = np.zeros(8, np.float32)
            if camera == 'kitti360_fisheye':
                intr[:7] = _host_f32(img_intrinsic_fisheye, 7)
                code = 3
            else:
                K = img_intrinsic_pinhole.detach().cpu().numpy() if isinstance(img_intrinsic_pinhole, torch.Tensor) \
                    else np.asarray(img_intrinsic_pinhole)
                K = np.asarray(K, dtype=np.float32)
                intr[:4] = [K[0, 0], K[1, 1], K[0, 2], K[1, 2]]
                code = 1
            cam = torch.from_numpy(np.concatenate([cam_xyz, A.reshape(-1), t0, t1, intr]).astype(np.float32)).to(dev)
            check(lib.dva_project_camera(ptr(xyz), ptr(cam), code, ptr(dist), ptr(x_proj), ptr(y_proj), ptr(keep),
                                         n, W, H, int(crop_top), int(crop_bottom), float(r_min), float(r_max),
                                         stream_ptr()), "dva_project_camera")
        else:
            raise ValueError(f"unknown camera '{camera}'")
    if img_mask is not None:  # field_of_view_cpu: img_mask[floor(x), floor(y)] (visibility.py:428-434)
        assert tuple(img_mask.shape) == (W, H), \
            f'Expected img_mask to be a torch.BoolTensor of shape img_size={img_size} but got size={img_mask.shape}.'
        xi = x_proj.floor().long().clamp(0, W - 1)
        yi = y_proj.floor().long().clamp(0, H - 1)
        keep = keep.bool() & img_mask.to(dev)[xi, yi]
    indices = torch.nonzero(keep, as_tuple=False).view(-1)
    return indices, dist[indices], x_proj[indices], y_proj[indices]


def _project_raw(xyz, camera, img_extrinsic, intr8, code):
    """x_proj, y_proj of every row of xyz (no filtering) through dva_project_camera."""
    lib = _lib.load()
    dev, n = xyz.device, xyz.shape[0]
    A, t0, t1 = _camera_transform(camera, img_extrinsic)
    cam = torch.from_numpy(np.concatenate([np.zeros(3, np.float32), A.reshape(-1), t0, t1, intr8])
                           .astype(np.float32)).to(dev)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    xp = torch.empty(n, dtype=torch.float64, device=dev)
    yp = torch.empty(n, dtype=torch.float64, device=dev)
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.dva_project_camera(ptr(xyz), ptr(cam), code, ptr(d), ptr(xp), ptr(yp), ptr(keep), n, 1 << 20,
                                     1 << 20, 0, 0, 0.0, 1e30, stream_ptr()), "dva_project_camera")
    return xp, yp


def fisheye_splat_boxes(x_proj, y_proj, xyz, img_extrinsic, img_intrinsic_fisheye, img_size=(1024, 512),
                        crop_top=0, crop_bottom=0, voxel=0.02, k_swell=1.0, d_swell=1000,
                        camera='kitti360_fisheye'):
    """fisheye_splat_cpu (visibility.py:876-953): the splat width is twice the image distance between
    a point and the projection of the top of its voxel (xyz + [0, 0, swell * voxel / 2]); NB the
    reference takes `dist = norm(xyz)` of the ABSOLUTE coordinates here (:900), reproduced."""
    require_cuda(x_proj, y_proj, xyz)
    lib = _lib.load()
    xyz = xyz.float().contiguous()
    m = xyz.shape[0]
    # norm_cpu: float32, squares summed left to right (a device-side reduction may associate otherwise)
    d = torch.sqrt((xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]) + xyz[:, 2] * xyz[:, 2])
    swell = 1 + k_swell * torch.exp(-d.double() / np.log(d_swell))          # float64, like numba
    top = xyz.clone()
    top[:, 2] += (swell * voxel / 2).float()                                # z_offset is float32
    intr = np.zeros(8, np.float32)
    intr[:7] = _host_f32(img_intrinsic_fisheye, 7)
    xt, yt = _project_raw(top, camera, img_extrinsic, intr, 3)
    width = 2 * torch.sqrt((x_proj.double() - xt) ** 2 + (y_proj.double() - yt) ** 2)
    splat = torch.empty((m, 4), dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(lib.dva_splat_boxes_from_width(ptr(x_proj.double().contiguous()), ptr(y_proj.double().contiguous()),
                                             ptr(width.contiguous()), ptr(splat), m, int(img_size[0]),
                                             int(img_size[1]), int(crop_top), int(crop_bottom), stream_ptr()),
              "dva_splat_boxes_from_width")
    return splat


def splat_boxes(x_proj, y_proj, dist, img_intrinsic_pinhole=None, img_size=(1024, 512), crop_top=0,
                crop_bottom=0, voxel=0.02, k_swell=1.0, d_swell=1000, camera='s3dis_equirectangular'):
    """[m,4] int32 (x_a, x_b, y_a, y_b) like *_splat_cpu (visibility.py:630-704, 761-827)."""
    require_cuda(x_proj, y_proj, dist)
    lib = _lib.load()
    m = x_proj.shape[0]
    splat = torch.empty((m, 4), dtype=torch.int32, device=x_proj.device)
    if camera == 's3dis_equirectangular':
        cam, fx, fy = 0, 0.0, 0.0
    elif camera in _PINHOLE_CAMERAS:
        cam = 1
        fx, fy = float(img_intrinsic_pinhole[0][0]), float(img_intrinsic_pinhole[1][1])
    else:
        raise NotImplementedError(f"camera='{camera}' has no CUDA splat kernel yet")
    D = ctypes.c_double
    with torch.cuda.device(x_proj.device):
        check(lib.dva_splat_boxes(ptr(x_proj.double().contiguous()), ptr(y_proj.double().contiguous()),
                                  ptr(dist.float().contiguous()), ptr(splat), m, int(img_size[0]),
                                  int(img_size[1]), int(crop_top), int(crop_bottom), D(float(voxel)),
                                  D(float(k_swell)), D(float(d_swell)), cam, D(fx), D(fy), stream_ptr()),
              "dva_splat_boxes")
    return splat


def visibility_from_splatting(x_proj, y_proj, dist, xyz=None, img_extrinsic=None, img_intrinsic_pinhole=None,
                              img_intrinsic_fisheye=None, img_size=(1024, 512), crop_top=0, crop_bottom=0,
                              voxel=0.1, k_swell=1.0, d_swell=1000, exact=False,
                              camera='s3dis_equirectangular', **kwargs):
    """Z-buffer visibility -> (indices, x_pix, y_pix) in the reference's order (row-major over the
    [x, y] winner map).  Ties go to the lowest point index; `exact` keeps splat centres only."""
    require_cuda(x_proj, y_proj, dist)
    assert x_proj.shape[0] == y_proj.shape[0] == dist.shape[0] > 0
    lib = _lib.load()
    dev = x_proj.device
    W, H = int(img_size[0]), int(img_size[1])
    Hc = H - int(crop_top) - int(crop_bottom)
    xp, yp = x_proj.double().contiguous(), y_proj.double().contiguous()
    d = dist.float().contiguous()
    m = d.shape[0]
    if camera == 'kitti360_fisheye':
        splat = fisheye_splat_boxes(xp, yp, xyz, img_extrinsic, img_intrinsic_fisheye, img_size, crop_top,
                                    crop_bottom, voxel, k_swell, d_swell, camera)
    else:
        splat = splat_boxes(xp, yp, d, img_intrinsic_pinhole, img_size, crop_top, crop_bottom, voxel, k_swell,
                            d_swell, camera)
    zbuf = torch.empty(W * Hc, dtype=torch.int64, device=dev)       # uint64 keys
    idx_map = torch.empty((W, Hc), dtype=torch.int64, device=dev)
    seen = torch.empty(m, dtype=torch.uint8, device=dev) if exact else None
    with torch.cuda.device(dev):
        check(lib.dva_zbuffer_splat(ptr(splat), ptr(d), ptr(xp), ptr(yp), ptr(zbuf), ptr(idx_map), ptr(seen), m,
                                    W, H, int(crop_top), int(crop_bottom), int(bool(exact)), stream_ptr()),
              "dva_zbuffer_splat")
    pix = torch.nonzero(idx_map >= 0, as_tuple=False)
    x_pix, y_pix = pix[:, 0], pix[:, 1]
    return idx_map[x_pix, y_pix], x_pix, y_pix + int(crop_top)


def postprocess_features(xyz_to_img, y_proj, dist, linearity, planarity, scattering, normals,
                         img_size=(1024, 512), r_max=30, r_min=0.5, **kwargs):
    """[n,F] viewing-condition features (visibility.py:1548-1582): normalised depth, linearity,
    planarity, scattering, |cos(view, normal)|, normalised pixel height."""
    features = []
    if dist is not None:
        # tensor / tensor: a CUDA division by a Python scalar is evaluated as a multiplication by its
        # reciprocal (1 ulp off the reference's CPU result, normalize_dist_cuda visibility.py:1503-1518)
        d = dist.float()
        features.append(((d - r_min) / torch.full_like(d, r_max + 1e-4)).float())
    for f in (linearity, planarity, scattering):
        if f is not None:
            features.append(f)
    if xyz_to_img is not None and dist is not None and normals is not None:
        u = (xyz_to_img / (dist + 1e-4).reshape((-1, 1))).float()
        p = u * normals.float()
        features.append(((p[:, 0] + p[:, 1]) + p[:, 2]).abs())           # torch CPU's sum order over 3 terms
    if y_proj is not None:
        features.append((y_proj / torch.full_like(y_proj, float(img_size[1]))).float())
    return torch.stack(features).t()


class VisibilityModel:
    """Same call contract as the reference's VisibilityModel (visibility.py:1677-1761)."""

    def __init__(self, img_size=(1024, 512), crop_top=0, crop_bottom=0, r_max=30, r_min=0.5,
                 camera='s3dis_equirectangular'):
        self.img_size = img_size
        self.crop_top = crop_top
        self.crop_bottom = crop_bottom
        self.r_max = r_max
        self.r_min = r_min
        self.camera = camera

    def _camera_projection(self, *args, **kwargs):
        return camera_projection(*args, **self.__dict__, **kwargs)

    def _visibility(self, *args, **kwargs):
        raise NotImplementedError

    def _postprocess_features(self, *args):
        return postprocess_features(*args, **self.__dict__)

    def __call__(self, xyz, img_xyz, linearity=None, planarity=None, scattering=None, normals=None, **kwargs):
        dev = xyz.device
        idx_1, dist, x_proj, y_proj = self._camera_projection(xyz, img_xyz, **kwargs)
        if x_proj.shape[0] == 0:
            e_long = torch.empty((0,), dtype=torch.long, device=dev)
            e_f = torch.empty((0,), dtype=torch.float, device=dev)
            return {'idx': e_long, 'x': e_long.clone(), 'y': e_long.clone(), 'depth': e_f, 'features': e_f.clone()}
        idx_2, x_pix, y_pix = self._visibility(x_proj, y_proj, dist, xyz[idx_1], **kwargs)
        idx = idx_1[idx_2]
        dist, y_proj = dist[idx_2], y_proj[idx_2]
        out = {'idx': idx, 'x': x_pix, 'y': y_pix, 'depth': dist}
        pick = lambda t: t[idx] if t is not None else None  # noqa: E731
        img_xyz_d = torch.as_tensor(img_xyz, dtype=xyz.dtype, device=dev)
        out['features'] = self._postprocess_features(xyz[idx] - img_xyz_d, y_proj, dist, pick(linearity),
                                                     pick(planarity), pick(scattering), pick(normals))
        return out

    def __repr__(self):
        return f"{self.__class__.__name__}({', '.join(f'{k}={v}' for k, v in self.__dict__.items())})"


class SplattingVisibility(VisibilityModel):
    """visibility.py:1764-1776."""

    def __init__(self, voxel=0.1, k_swell=1.0, d_swell=1000, exact=False, **kwargs):
        super().__init__(**kwargs)
        self.voxel = voxel
        self.k_swell = k_swell
        self.d_swell = d_swell
        self.exact = exact

    def _visibility(self, x_proj, y_proj, dist, xyz, **kwargs):
        return visibility_from_splatting(x_proj, y_proj, dist, xyz, **self.__dict__, **kwargs)
