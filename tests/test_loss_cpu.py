"""CPU: the loss-epilogue oracle (oracle/loss_oracle.py) against tests/golden/loss.npz, which holds values and autograd gradients of
the reference's OWN `_unpack_imgs` and `img2mse` (core/train/trainers/occnerf/trainer.py:24,31-41, executed from the source text)."""
import os

import numpy as np
import torch

from oracle import loss_oracle as L

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss.npz")


def test_oracle_reproduces_the_reference_loss():
    g = np.load(GOLDEN)
    rgbs = torch.from_numpy(g["rgbs"]).requires_grad_(True)
    comp = torch.from_numpy(g["comp"]).requires_grad_(True)
    total, imgs = L.loss(rgbs, comp, torch.from_numpy(g["patch_masks"]), torch.from_numpy(g["bgcolor"]), torch.from_numpy(g["targets"]),
                         g["div"].tolist(), float(g["w_mse"]), float(g["w_comp"]))
    total.backward()
    assert np.array_equal(imgs.detach().numpy(), g["patch_imgs"])
    assert abs(float(total) - float(g["loss"])) < 1e-6
    assert np.abs(rgbs.grad.numpy() - g["g_rgbs"]).max() < 1e-9 and np.abs(comp.grad.numpy() - g["g_comp"]).max() < 1e-9
