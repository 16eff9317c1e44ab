"""CPU restatement of NeighborhoodBasedMappingFeatures (TEST INFRASTRUCTURE ONLY -- see
oracle/__init__.py; never imported by the product path).

Follows torch_points3d/core/data_transform/multimodal/image.py:483-612 (KeOps branch).  Pinned on
tests/golden/neighborhood_features.npz, produced by running the reference class itself with a
dense stand-in for the KeOps LazyTensor (oracle/ref_loader.py: exact search, ties by index).
"""
import numpy as np


def knn_bruteforce(pos, k, block=512):
    """k nearest neighbours (self included) of every point; squared distances
    (dx*dx + dy*dy) + dz*dz in float32, ascending (dist, index)  -- image.py:504-514.
    Returns (neighbors [N,k] int64, dist2 [N,k] float32)."""
    p = np.asarray(pos, dtype=np.float32)
    n = p.shape[0]
    nbr = np.empty((n, k), dtype=np.int64)
    d2o = np.empty((n, k), dtype=np.float32)
    for s in range(0, n, block):
        q = p[s:s + block]
        dx = q[:, None, 0] - p[None, :, 0]
        dy = q[:, None, 1] - p[None, :, 1]
        dz = q[:, None, 2] - p[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        idx = np.argsort(d2, axis=1, kind="stable")[:, :k]
        nbr[s:s + block] = idx
        d2o[s:s + block] = np.take_along_axis(d2, idx, axis=1)
    return nbr, d2o


def neighborhood_features(pos, neighbors, pointers, images, k_list, voxel=1, density=True, occlusion=True):
    """[V, nk*(density+occlusion)] float32: densities for ascending k, then occlusions."""
    f32 = np.float32
    p = np.asarray(pos, dtype=f32)
    nbr = np.asarray(neighbors)
    ptr = np.asarray(pointers).astype(np.int64)
    img = np.asarray(images).astype(np.int64)
    counts = ptr[1:] - ptr[:-1]
    n = p.shape[0]
    view_point = np.repeat(np.arange(n), counts)
    cols = []
    k_list = sorted(k_list)
    if density:
        for k in k_list:
            d = p - p[nbr[:, k - 1]]                                    # :527
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            with np.errstate(divide="ignore", invalid="ignore"):
                v_sphere = f32(3.1416) * d2                             # :532
                den = (f32(k + 1) / v_sphere) / f32(1 / voxel ** 2)     # :533-534
            den[np.isnan(den)] = 1                                      # :537
            cols.append(den.astype(f32)[view_point])                    # :546
    if occlusion:
        n_img = int(img.max()) + 1 if img.size else 0
        seen_table = np.zeros((n, n_img), dtype=bool)                   # :567-569
        seen_table[view_point, img] = True
        for k in k_list:
            seen = np.ones(img.shape[0], dtype=f32)                     # :575
            for i in range(k):
                seen += seen_table[nbr[view_point, i], img]             # :577-580
            cols.append(seen / f32(k + 1))                              # :584
    return np.stack(cols, axis=1).astype(f32)
