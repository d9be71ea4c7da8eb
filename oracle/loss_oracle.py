"""ORACLE (test infrastructure): torch-CPU restatement of the reference's loss epilogue without the LPIPS term,
core/train/trainers/occnerf/trainer.py:31-41 (_unpack_imgs), :24 (img2mse), :135-147 and :172-189 (get_loss: weighted sum of the image
loss and mean(comp_loss)).  Pinned by tests/golden/loss.npz, written by oracle/make_golden_loss.py from the reference's own text."""
from __future__ import annotations

import torch


def unpack_imgs(rgbs, patch_masks, bgcolor, targets, div_indices):
    n_patch = len(div_indices) - 1
    imgs = bgcolor.expand(targets.shape).clone()                                     # trainer.py:36
    for i in range(n_patch):
        imgs[i, patch_masks[i]] = rgbs[div_indices[i]:div_indices[i + 1]]            # :38-39
    return imgs


def loss(rgbs, comp_loss, patch_masks, bgcolor, targets, div_indices, w_mse, w_comp):
    imgs = unpack_imgs(rgbs, patch_masks, bgcolor, targets, div_indices)
    mse = torch.mean((imgs - targets) ** 2)                                          # img2mse, :24 / :96-97
    comp = torch.mean(comp_loss)                                                     # :172-175
    return w_mse * mse + w_comp * comp, imgs
