"""Writes tests/golden/patches.npz with the reference's OWN patch-selection code: the two methods of
core/data/occnerf/train.py (get_patch_ray_indices :167-222, _get_patch_ray_indices :225-273) are compiled from the reference's
source text (the Dataset class itself needs cv2, the config system and a dataset on disk) and run under a seeded np.random whose
draws are recorded, so that the restatement and the CUDA kernel can be fed the same draws.  Run in the build container only."""
import ast
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OCCNERF_REFERENCE", "/root/reference")


def _reference_methods(ratio):
    src = open(os.path.join(REF, "core", "data", "occnerf", "train.py")).read()
    fns = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name in ("get_patch_ray_indices", "_get_patch_ray_indices"):
            fns[node.name] = node
    if not hasattr(np, "bool"):
        np.bool = bool                                     # the reference predates numpy 1.24
    cfg = types.SimpleNamespace(patch=types.SimpleNamespace(sample_subject_ratio=ratio))
    ns = {"np": np, "cfg": cfg}
    for node in fns.values():
        exec(compile(ast.Module([node], []), "train.py", "exec"), ns)
    obj = types.SimpleNamespace()
    obj._get_patch_ray_indices = types.MethodType(ns["_get_patch_ray_indices"], obj)
    return types.MethodType(ns["get_patch_ray_indices"], obj)


class _Recorder:
    """np.random stand-in that forwards to a seeded RandomState and records what the reference drew."""

    def __init__(self, seed):
        self.rs, self.u, self.choice_n, self.choice = np.random.RandomState(seed), [], [], []

    def rand(self, *a):
        v = self.rs.rand(*a)
        self.u.append(float(v[0]))
        return v

    def choice(self, n, size, replace):
        raise NotImplementedError


def case(H, W, patch, n_patch, seed, ratio):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    # a box-shaped "bbox" region touching two image borders, a blob-shaped subject inside it, holes in the ray mask
    bbox = (yy >= 0) & (yy < int(0.8 * H)) & (xx >= int(0.1 * W)) & (xx < W)
    subject = ((yy - 0.4 * H) ** 2 / (0.3 * H) ** 2 + (xx - 0.55 * W) ** 2 / (0.2 * W) ** 2) < 1.0
    subject &= bbox
    ray_mask = bbox.copy()
    ray_mask &= rng.random((H, W)) > 0.02                  # (the reference's ray_mask is the bbox hit mask; holes exercise the ranks)
    fn = _reference_methods(ratio)
    rec = _Recorder(seed)
    draws = []
    real = np.random

    class R:                                               # what the reference code sees as np.random
        @staticmethod
        def rand(*a):
            return rec.rand(*a)

        @staticmethod
        def choice(n, size, replace):
            v = rec.rs.choice(n, size=size, replace=replace)
            draws.append(int(v[0]))
            return v

    np.random = R
    try:
        select_inds, info, div = fn(n_patch, ray_mask.reshape(-1), subject, bbox, patch, H, W)
    finally:
        np.random = real
    return dict(H=H, W=W, patch=patch, ratio=ratio, ray_mask=ray_mask.reshape(-1), subject_mask=subject, bbox_mask=bbox,
                u=np.array(rec.u), select_idx=np.array(draws), select_inds=select_inds, patch_masks=info["mask"],
                xy_min=info["xy_min"], xy_max=info["xy_max"], patch_div_indices=div)


def main():
    out = {}
    for name, kw in {"a": dict(H=96, W=128, patch=32, n_patch=6, seed=3, ratio=0.8),
                     "b": dict(H=70, W=50, patch=20, n_patch=9, seed=11, ratio=0.5)}.items():
        for k, v in case(**kw).items():
            out[f"{name}_{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "patches.npz")
    np.savez_compressed(path, **out)
    print(path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith(("select_inds", "patch_div_indices"))})


if __name__ == "__main__":
    main()
