"""CPU: host-side logic of the sharded view render (occnerf_b200/render.py): per-rank frames painted from contiguous
ray ranges merge into exactly the single-rank frame, for any world size (incl. more ranks than rays)."""
import os

import numpy as np
import pytest

from occnerf_b200 import distributed as D
from occnerf_b200 import render
from oracle import image_oracle as IO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_merge_frames_equals_single_rank(world):
    g = np.load(os.path.join(GOLDEN, "image_unpack.npz"))
    H, W, mask, bg = int(g["H"]), int(g["W"]), g["ray_mask"], g["bgcolor"] / 255.
    n = int(mask.sum())
    pix = np.nonzero(mask)[0]
    frames, spans = [], []
    for r in range(world):
        b, e = D.shard_range(n, r, world)
        own = np.zeros_like(mask)
        own[pix[b:e]] = True
        frames.append(IO.unpack(W, H, own, bg, g["rgb"][b:e], g["alpha"][b:e]))   # what rank r paints
        spans.append((b, e))
    rgb8, alpha8 = render.merge_frames(frames, mask, spans)
    assert np.array_equal(rgb8, g["rgb_image"]) and np.array_equal(alpha8, g["alpha_image"][..., 0])


def test_merge_frames_with_more_ranks_than_rays():
    mask = np.zeros(16, bool)
    mask[[3, 9]] = True
    rgb, alpha = np.array([[0.2, 0.4, 0.6], [1.0, 0.0, 0.5]], np.float32), np.array([0.5, 1.0], np.float32)
    full = IO.unpack(4, 4, mask, [0, 0, 0], rgb, alpha)
    frames, spans = [], []
    for r in range(4):
        b, e = D.shard_range(2, r, 4)
        own = np.zeros_like(mask)
        own[np.nonzero(mask)[0][b:e]] = True
        frames.append(IO.unpack(4, 4, own, [0, 0, 0], rgb[b:e], alpha[b:e]))
        spans.append((b, e))
    got = render.merge_frames(frames, mask, spans)
    assert np.array_equal(got[0], full[0]) and np.array_equal(got[1], full[1])
