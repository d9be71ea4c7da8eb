"""GPU: occnerf_sample_patches (csrc/patches.cu) bit-exact against the fixture written by the reference's own patch-selection code
(tests/golden/patches.npz) and against the oracle on a 512 x 512 case with the rays of the benchmark's camera, incl. the gather of the
selected rays and the out-of-range status."""
import os

import numpy as np
import pytest
import torch

from occnerf_b200 import ops
from oracle import patch_oracle as P
from tests.helpers import dev

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "patches.npz")


def _run(ray_mask, subject, bbox, H, W, patch, use, idx, rays=None):
    d = dev()
    out = ops.sample_patches(torch.from_numpy(ray_mask.reshape(-1)).to(d), torch.from_numpy(subject.reshape(-1)).to(d),
                             torch.from_numpy(bbox.reshape(-1)).to(d), H, W, patch, use, idx, rays=rays)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}


def test_reference_fixture_bit_exact():
    g = np.load(GOLDEN)
    for n in "ab":
        H, W, patch = int(g[f"{n}_H"]), int(g[f"{n}_W"]), int(g[f"{n}_patch"])
        use = g[f"{n}_u"] < float(g[f"{n}_ratio"])
        o = _run(g[f"{n}_ray_mask"], g[f"{n}_subject_mask"], g[f"{n}_bbox_mask"], H, W, patch, use, g[f"{n}_select_idx"])
        div = g[f"{n}_patch_div_indices"]
        assert o["status"][0] == 0
        assert np.array_equal(o["patch_div"], div)
        assert np.array_equal(o["select_inds"][:div[-1]], g[f"{n}_select_inds"])
        assert np.array_equal(o["patch_masks"].astype(bool), g[f"{n}_patch_masks"])
        assert np.array_equal(o["xy_min"], g[f"{n}_xy_min"]) and np.array_equal(o["xy_max"], g[f"{n}_xy_max"])


def test_against_oracle_512_with_ray_gather():
    rng = np.random.default_rng(0)
    H = W = 512
    yy, xx = np.mgrid[0:H, 0:W]
    bbox = (np.abs(yy - 250) < 200) & (np.abs(xx - 260) < 120)
    subject = ((yy - 250) ** 2 / 170.0 ** 2 + (xx - 260) ** 2 / 60.0 ** 2) < 1.0
    ray_mask = bbox & (rng.random((H, W)) > 0.01)
    n_rays = int(ray_mask.sum())
    rays = torch.from_numpy(rng.standard_normal((n_rays, 8)).astype(np.float32)).to(dev())
    use, idx = ops.draw_patch_randoms(6, 0.8, int(subject.sum()), int((bbox & ~subject).sum()), rs=np.random.RandomState(5))
    want = P.sample_patches(ray_mask, subject, bbox, 32, H, W, use, idx)
    o = _run(ray_mask, subject, bbox, H, W, 32, use, idx, rays=rays)
    n = int(want[4][-1])
    assert o["status"][0] == 0 and np.array_equal(o["patch_div"], want[4]) and np.array_equal(o["select_inds"][:n], want[0])
    assert np.array_equal(o["patch_masks"].astype(bool), want[1]) and np.array_equal(o["xy_min"], want[2]) and np.array_equal(o["xy_max"], want[3])
    assert np.array_equal(o["rays"][:n], rays.cpu().numpy()[want[0]])
    # a draw beyond the candidate list (np.random.choice could not produce it) is reported
    bad = _run(ray_mask, subject, bbox, H, W, 32, np.array([True]), np.array([int(subject.sum())]))
    assert bad["status"][0] == 1
