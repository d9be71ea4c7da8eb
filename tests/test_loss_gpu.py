"""GPU: occnerf_patch_loss (csrc/loss.cu) against the fixture of the reference's own `_unpack_imgs` + `img2mse` + mean(comp_loss)
(tests/golden/loss.npz): patch images bit-exact, loss and both gradients to fp32 rounding; and at the benchmark's shape
(6 x 32 x 32 fully covered patches, 786 432 completeness terms) against the oracle."""
import os

import numpy as np
import pytest
import torch

from occnerf_b200 import ops
from oracle import loss_oracle as L
from tests.helpers import dev

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss.npz")


def test_reference_fixture():
    g, d = np.load(GOLDEN), dev()
    rgbs = torch.from_numpy(g["rgbs"]).to(d).requires_grad_(True)
    comp = torch.from_numpy(g["comp"]).to(d).requires_grad_(True)
    loss, imgs, parts = ops.patch_loss(rgbs, comp, torch.from_numpy(g["patch_masks"]).to(d), torch.from_numpy(g["div"]).to(d),
                                       torch.from_numpy(g["bgcolor"]).to(d), torch.from_numpy(g["targets"]).to(d), float(g["w_mse"]), float(g["w_comp"]))
    (3.0 * loss).backward()
    assert np.array_equal(imgs.detach().cpu().numpy(), g["patch_imgs"])
    assert abs(float(loss) - float(g["loss"])) < 2e-6 * abs(float(g["loss"]))
    assert np.abs(rgbs.grad.cpu().numpy() - 3.0 * g["g_rgbs"]).max() < 1e-9
    assert np.abs(comp.grad.cpu().numpy() - 3.0 * g["g_comp"]).max() < 1e-9
    assert abs(float(parts.sum()) - float(loss)) < 1e-6
    # a consumer of the unpacked images (the perceptual term) gets its gradient back through the scatter
    rgbs.grad = None
    w = torch.from_numpy(g["targets"]).to(d)
    loss2, imgs2, _ = ops.patch_loss(rgbs, None, torch.from_numpy(g["patch_masks"]).to(d), torch.from_numpy(g["div"]).to(d),
                                     torch.from_numpy(g["bgcolor"]).to(d), torch.from_numpy(g["targets"]).to(d), 0.0, 0.0)
    (imgs2 * w).sum().backward()
    a = torch.from_numpy(g["rgbs"]).requires_grad_(True)
    (L.unpack_imgs(a, torch.from_numpy(g["patch_masks"]), torch.from_numpy(g["bgcolor"]), torch.from_numpy(g["targets"]), g["div"].tolist())
     * torch.from_numpy(g["targets"])).sum().backward()
    assert float((rgbs.grad.cpu() - a.grad).abs().max()) < 1e-7


def test_bench_shape_against_oracle():
    d = dev()
    gen = torch.Generator().manual_seed(1)
    N, P, S = 6, 32, 128
    rgb = torch.rand(N * P * P, 3, generator=gen)
    comp = torch.rand(N * P * P, S, generator=gen)
    targets = torch.rand(N, P, P, 3, generator=gen)
    masks = torch.ones(N, P, P, dtype=torch.bool)
    div = [i * P * P for i in range(N + 1)]
    bg = torch.tensor([0.0, 0.0, 0.0])
    a, b = rgb.clone().requires_grad_(True), comp.clone().requires_grad_(True)
    want, _ = L.loss(a, b, masks, bg, targets, div, 0.2, 1.0)
    want.backward()
    x, y = rgb.to(d).requires_grad_(True), comp.to(d).requires_grad_(True)
    got, _, _ = ops.patch_loss(x, y, masks.to(d), torch.tensor(div, device=d), bg.to(d), targets.to(d), 0.2, 1.0)
    got.backward()
    assert abs(float(got) - float(want)) < 1e-6 * abs(float(want))
    assert float((x.grad.cpu() - a.grad).abs().max()) < 1e-10 and float((y.grad.cpu() - b.grad).abs().max()) < 1e-12
