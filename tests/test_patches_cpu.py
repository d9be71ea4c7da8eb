"""CPU: the patch-selection oracle (oracle/patch_oracle.py) against tests/golden/patches.npz, which holds the outputs of the
reference's OWN get_patch_ray_indices / _get_patch_ray_indices (core/data/occnerf/train.py:167-273, executed from the reference's
source text by oracle/make_golden_patches.py) together with the random draws the reference took."""
import os

import numpy as np

from oracle import patch_oracle as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "patches.npz")


def test_oracle_reproduces_the_reference_patches():
    g = np.load(GOLDEN)
    for n in "ab":
        use = g[f"{n}_u"] < float(g[f"{n}_ratio"])
        got = P.sample_patches(g[f"{n}_ray_mask"], g[f"{n}_subject_mask"], g[f"{n}_bbox_mask"], int(g[f"{n}_patch"]), int(g[f"{n}_H"]),
                               int(g[f"{n}_W"]), use, g[f"{n}_select_idx"])
        want = (g[f"{n}_select_inds"], g[f"{n}_patch_masks"], g[f"{n}_xy_min"], g[f"{n}_xy_max"], g[f"{n}_patch_div_indices"])
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        assert use.any() and not use.all(), "both candidate regions must occur in the fixture"


def test_draws_follow_the_reference_order():
    """draw_patch_randoms consumes numpy's RandomState exactly as the reference does (rand, then choice without replacement)."""
    from occnerf_b200 import ops
    g = np.load(GOLDEN)
    for n, seed in (("a", 3), ("b", 11)):
        subject, bbox = g[f"{n}_subject_mask"], g[f"{n}_bbox_mask"]
        use, idx = ops.draw_patch_randoms(len(g[f"{n}_u"]), float(g[f"{n}_ratio"]), int(subject.sum()), int((bbox & ~subject).sum()),
                                          rs=np.random.RandomState(seed))
        assert np.array_equal(use, g[f"{n}_u"] < float(g[f"{n}_ratio"])) and np.array_equal(idx, g[f"{n}_select_idx"])
