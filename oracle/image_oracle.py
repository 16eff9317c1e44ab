"""CPU restatement of the reference's bilinear feature-map sampling (TEST INFRASTRUCTURE ONLY --
see oracle/__init__.py; never imported by the product path).

Pinned on tests/golden/sparse_interpolation.npz, which oracle/make_golden.py produced by running
the reference's own `sparse_interpolation` (torch_points3d/core/multimodal/image.py:105-170) the
way `get_mapped_features(interpolate=True)` calls it (image.py:1278-1283).
"""
import numpy as np


def sparse_interpolation_pixels(x, pix, batch, mapping_size):
    """x [B,C,h,w] float32; pix [P,2] integer (x, y) at the mapping resolution `mapping_size` =
    (W, H); batch [P].  Returns [P,C] float32.  Every step in float32, in the reference's order:
      coords = pix / (resolution - 1), swapped to (row, col)          image.py:1280-1281
      p = coords * (h, w) + 0.5 in the 1-px replicate-padded frame    image.py:133, 143
      corners floor(p), floor(p + 1)                                  image.py:149-156
      weight of a corner = |prod(p - opposite corner)|                image.py:159-162
      out = w_tl f_tl + w_tr f_tr + w_bl f_bl + w_br f_br             image.py:164-167
    """
    f32 = np.float32
    x = np.asarray(x, dtype=f32)
    B, C, h, w = x.shape
    W, H = mapping_size
    px = np.asarray(pix)[:, 0].astype(f32)
    py = np.asarray(pix)[:, 1].astype(f32)
    cy = py / f32(H - 1)
    cx = px / f32(W - 1)
    p0 = cy * f32(h) + f32(0.5)
    p1 = cx * f32(w) + f32(0.5)
    top, bottom = np.floor(p0), np.floor(p0 + f32(1))
    left, right = np.floor(p1), np.floor(p1 + f32(1))
    w_tl = np.abs((p0 - bottom) * (p1 - right))
    w_tr = np.abs((p0 - bottom) * (p1 - left))
    w_bl = np.abs((p0 - top) * (p1 - right))
    w_br = np.abs((p0 - top) * (p1 - left))
    padded = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")
    b = np.asarray(batch).astype(np.int64)

    def at(r, c):
        return padded[b, :, r.astype(np.int64), c.astype(np.int64)]

    out = w_tl[:, None] * at(top, left) + w_tr[:, None] * at(top, right)
    out = out + w_bl[:, None] * at(bottom, left)
    out = out + w_br[:, None] * at(bottom, right)
    return out.astype(f32)
