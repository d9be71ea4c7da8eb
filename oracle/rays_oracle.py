"""ORACLE (test infrastructure -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it).

numpy restatement of the ray set-up in front of the rendering path:
  * `pixel_rays`      follows core/utils/camera_util.py:133-160 (get_rays_from_KRT)
  * `box_near_far`    follows core/utils/camera_util.py:163-212 (rays_intersect_3d_bbox, use_mask=True)
  * `frame_rays`      follows the masking every dataset does next (core/data/occnerf/freeview.py:208-219,
                      train.py:440-461): valid rays in pixel order, float32, packed as (o3, d3, near, far)
Pinned against the reference's own functions by tests/golden/rays_*.npz (written by oracle/make_golden_rays.py,
which imports camera_util.py from /root/reference in the build container); see tests/test_oracle_golden.py.
dtypes are left to numpy exactly as the reference leaves them: a float32 K gives a float32 `pixel_camera`.
"""
from __future__ import annotations

import numpy as np


def pixel_rays(H: int, W: int, K: np.ndarray, R: np.ndarray, T: np.ndarray):
    origin = -(R.T @ T).reshape(3)                                              # :148
    cols = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)                 # :150-152  i = column, j = row
    rows = np.arange(H, dtype=np.float32)[:, None].repeat(W, 1)
    homog = np.concatenate([cols[..., None], rows[..., None], np.ones((H, W, 1), np.float32)], -1)
    # np.dot on the 3-D operand, as the reference calls it: numpy evaluates an N-D x 2-D dot one output element at a time
    # through the BLAS dot routine, whose float32 rounding differs from what `@` / a 2-D gemm give (1 ulp in rays_d)
    cam = np.dot(homog, np.linalg.inv(K).T)                                     # :154
    world = np.dot(cam - T.reshape(3), R)                                       # :155
    dirs = world - origin                                                       # :157
    return np.broadcast_to(origin, dirs.shape), dirs


def box_near_far(bbox_min, bbox_max, ray_o: np.ndarray, ray_d: np.ndarray):
    """ray_o, ray_d [P,3]; ray_d is clamped IN PLACE like the reference (:183).  -> near [n], far [n], mask [P]."""
    lo = np.asarray(bbox_min).astype(np.float64) + -0.01                        # :180 (float32 box + float64 margin)
    hi = np.asarray(bbox_max).astype(np.float64) + 0.01
    small = np.abs(ray_d) < 1e-5                                                # :183
    ray_d[small] = 1e-5
    planes = np.concatenate([lo, hi])                                           # order (min xyz, max xyz)  :181,184
    axis = np.array([0, 1, 2, 0, 1, 2])
    steps = (planes[None, :] - ray_o[:, axis]) / ray_d[:, axis]                 # [P,6]
    pts = steps[:, :, None] * ray_d[:, None, :] + ray_o[:, None, :]             # :186  [P,6,3]
    tol = 1e-6                                                                  # :189
    inside = np.ones(steps.shape, bool)
    for a in range(3):
        inside &= (pts[..., a] >= lo[a] - tol) & (pts[..., a] <= hi[a] + tol)   # :190-195
    mask = inside.sum(1) == 2                                                   # :197
    seg = pts[mask][inside[mask]].reshape(-1, 2, 3)                             # :201  first / second plane hit in plane order
    o, d = ray_o[mask], ray_d[mask]
    length = np.sqrt((d * d).sum(1))
    t0 = np.sqrt(((seg[:, 0] - o) ** 2).sum(1)) / length                        # :207-208
    t1 = np.sqrt(((seg[:, 1] - o) ** 2).sum(1)) / length
    return np.minimum(t0, t1), np.maximum(t0, t1), mask


def frame_rays(H: int, W: int, K, R, T, bbox_min, bbox_max):
    """-> packed [n,8] float32 (o, d, near, far), mask [H*W] bool, pixel index [n] of each valid ray."""
    o, d = pixel_rays(H, W, K, R, T)
    o = np.ascontiguousarray(o.reshape(-1, 3))
    d = np.ascontiguousarray(d.reshape(-1, 3))
    near, far, mask = box_near_far(bbox_min, bbox_max, o, d)
    packed = np.concatenate([o[mask], d[mask], near[:, None], far[:, None]], 1).astype(np.float32)
    return packed, mask, np.nonzero(mask)[0].astype(np.int32)
