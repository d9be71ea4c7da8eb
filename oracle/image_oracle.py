"""ORACLE (test infrastructure): numpy restatement of the reference's image assembly, run.py:39-66
(unpack_alpha_map, unpack_to_image) with core/utils/image_util.py:19-20 (to_8b_image = uint8(255. * clip(x, 0, 1))).
Pinned by tests/golden/image_unpack.npz, written by oracle/make_golden_rays.py with the reference's own functions."""
from __future__ import annotations

import numpy as np


def eight_bit(x: np.ndarray) -> np.ndarray:
    return (np.float32(255.0) * np.minimum(np.maximum(x.astype(np.float32), np.float32(0)), np.float32(1))).astype(np.uint8)


def unpack(W: int, H: int, ray_mask: np.ndarray, bgcolor_01, rgb: np.ndarray, alpha: np.ndarray):
    """-> (rgb8 [H,W,3], alpha8 [H,W]); alpha8 is one channel of the reference's three identical ones."""
    frame = np.empty((H * W, 3), np.float32)
    frame[:] = np.asarray(bgcolor_01, np.float64).astype(np.float32)          # np.full(..., dtype='float32')  run.py:49
    frame[ray_mask] = rgb                                                       # :52
    cover = np.zeros(H * W, np.float32)                                         # :40-41
    cover[ray_mask] = alpha
    return eight_bit(frame).reshape(H, W, 3), eight_bit(cover).reshape(H, W)
