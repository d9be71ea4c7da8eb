"""ORACLE (test infrastructure): numpy restatement of the reference's training-patch selection,
core/data/occnerf/train.py:167-222 (get_patch_ray_indices) and :225-273 (_get_patch_ray_indices), with the two random draws of
every patch INJECTED instead of taken from np.random (the reference draws `np.random.rand(1)[0] < sample_subject_ratio` and
`np.random.choice(n_candidates, size=[1], replace=False)[0]`; the caller keeps drawing them with numpy, in that order).
Pinned by tests/golden/patches.npz, which oracle/make_golden_patches.py writes by executing the reference's own function text
under a recording np.random."""
from __future__ import annotations

import numpy as np


def sample_patches(ray_mask, subject_mask, bbox_mask, patch_size, H, W, use_subject, select_idx):
    """ray_mask [H*W] bool, subject_mask / bbox_mask [H,W] bool; use_subject [n] bool, select_idx [n] int (index into the row-major list
    of candidate pixels) -> (select_inds [sum], patch_masks [n,P,P] bool, xy_min [n,2], xy_max [n,2], patch_div_indices [n+1])."""
    ray_mask = np.asarray(ray_mask, bool).reshape(-1)
    bbox_ex = np.bitwise_and(bbox_mask, np.bitwise_not(subject_mask))                       # train.py:183-186
    masked_indices = np.cumsum(ray_mask) - 1                                                # :262
    inds, masks, mins, maxs, div = [], [], [], [], [0]
    for use, sel in zip(use_subject, select_idx):
        cand = subject_mask if use else bbox_ex                                            # :199-202
        ys, xs = np.where(cand)                                                            # :238
        cx, cy = xs[sel], ys[sel]                                                          # :241-244
        half = patch_size // 2
        x_min = int(np.clip(cx - half, 0, W - patch_size)); y_min = int(np.clip(cy - half, 0, H - patch_size))   # :247-254
        sel_mask = np.zeros((H, W), bool)
        sel_mask[y_min:y_min + patch_size, x_min:x_min + patch_size] = True                # :256-257
        inter = np.bitwise_and(sel_mask.reshape(-1), ray_mask)                             # :263-264
        inds.append(masked_indices[np.where(inter)])                                       # :265-268
        masks.append(inter.reshape(H, W)[y_min:y_min + patch_size, x_min:x_min + patch_size])
        mins.append([x_min, y_min]); maxs.append([x_min + patch_size, y_min + patch_size])
        div.append(div[-1] + len(inds[-1]))
    return np.concatenate(inds), np.stack(masks), np.array(mins), np.array(maxs), np.array(div)
