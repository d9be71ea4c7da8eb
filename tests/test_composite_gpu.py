"""GPU: warp-scan compositing (occnerf_composite_*) against the golden fixtures and the oracle (torch autograd on CPU)."""
import numpy as np
import pytest
import torch

from occnerf_b200 import ops
from oracle import make_golden, occnerf_oracle as O
from tests.helpers import dev, load_case, maxabs, normwise_close, report

pytestmark = pytest.mark.gpu


def _cuda(*ts):
    return [t.to(dev()).contiguous() for t in ts]


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_forward_against_reference_golden(name):
    sub, w, fr, vol, t_rand, rk, g = load_case(name)
    raw, mask, z = torch.from_numpy(g["raw"]), torch.from_numpy(g["mask"]), torch.from_numpy(g["z"])
    rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1)
    rgb, acc, depth, term, wts, comp = ops.composite_forward(*_cuda(raw, mask, z, rays, fr.bgcolor), want_weights=True, want_comp=True)
    errs = dict(rgb=maxabs(rgb, g["rgb"]), alpha=maxabs(acc, g["alpha"]), depth=maxabs(depth, g["depth"]), weights=maxabs(wts, g["weights"]))
    report(f"composite_golden[{name}]", **errs)
    assert max(errs[k] for k in ('rgb', 'alpha', 'weights')) < 2e-6 and errs['depth'] < 5e-6   # depth is O(6)
    assert np.array_equal(term.cpu().numpy().astype(np.int32), g["term"])
    if rk["training"]:
        assert maxabs(comp, g["comp_loss"]) < 1e-6


def _random_case(N, S, seed, bg=(30.0, 120.0, 255.0)):
    gen = torch.Generator().manual_seed(seed)
    raw = torch.randn(N, S, 5, generator=gen) * 2.0
    raw[..., 3] = raw[..., 3] * 3.0 + 1.0
    raw[..., 4] = torch.rand(N, S, generator=gen) * 0.8 - 0.3
    mask = torch.rand(N, S, generator=gen)
    mask[torch.rand(N, S, generator=gen) < 0.3] = 0.0
    near = torch.rand(N, 1, generator=gen) * 0.5 + 0.2
    z = torch.sort(near + torch.rand(N, S, generator=gen) * 0.05, dim=1)[0]
    rays = torch.cat([torch.randn(N, 3, generator=gen), torch.randn(N, 3, generator=gen) * 30.0, near, near + 2], -1)
    if N > 2:
        mask[1] = 0.0                      # a ray that hits nothing
        raw[2, :, 3] = 25.0                # opaque from the first sample, softplus threshold branch
        mask[2] = 1.0
    return raw, mask, z, rays, torch.tensor(bg)


@pytest.mark.parametrize("N,S", [(257, 128), (33, 64), (5, 256), (19, 100), (1, 32), (3, 7)])
def test_forward_backward_against_oracle(N, S):
    raw, mask, z, rays, bg = _random_case(N, S, seed=N * 1000 + S)
    rawr, maskr = raw.clone().requires_grad_(True), mask.clone().requires_grad_(True)
    rgb_o, acc_o, depth_o, term_o, w_o = O.composite(rawr, maskr, z, rays[:, 3:6], bg)
    comp_o = O.completeness_term(rawr)
    gen = torch.Generator().manual_seed(7)
    g_rgb, g_acc, g_depth, g_comp = torch.randn(N, 3, generator=gen), torch.randn(N, generator=gen), torch.randn(N, generator=gen), torch.randn(N, S, generator=gen)
    ((rgb_o * g_rgb).sum() + (acc_o * g_acc).sum() + (depth_o * g_depth).sum() + (comp_o * g_comp).sum()).backward()
    c = _cuda(raw, mask, z, rays, bg)
    rgb, acc, depth, term, wts, comp = ops.composite_forward(*c, want_weights=True, want_comp=True)
    assert maxabs(rgb, rgb_o) < 5e-6 and maxabs(acc, acc_o) < 5e-6 and maxabs(depth, depth_o) < 5e-6 and maxabs(wts, w_o) < 2e-6
    assert maxabs(comp, comp_o) < 1e-5
    assert torch.equal(term.cpu(), term_o)
    g_raw, g_mask = ops.composite_backward(*c, *_cuda(g_rgb, g_acc, g_depth, g_comp))
    e1 = maxabs(g_raw, rawr.grad) / float(rawr.grad.abs().max())
    e2 = maxabs(g_mask, maskr.grad) / float(maskr.grad.abs().max())
    report(f"composite_bwd[{N}x{S}]", g_raw_rel=e1, g_mask_rel=e2)
    assert normwise_close(g_raw.cpu().numpy(), rawr.grad.numpy(), 2e-5)
    assert normwise_close(g_mask.cpu().numpy(), maskr.grad.numpy(), 2e-5)


def test_empty_and_properties_at_full_size():
    d = dev()
    out = ops.composite_forward(torch.zeros(0, 128, 5, device=d), torch.zeros(0, 128, device=d), torch.zeros(0, 128, device=d),
                                torch.zeros(0, 8, device=d), torch.zeros(3, device=d))
    assert out[0].shape == (0, 3)
    # BASELINE config 2 size: 6144 rays x 128 samples; size-independent invariants
    raw, mask, z, rays, bg = _random_case(6144, 128, seed=3, bg=(255.0, 255.0, 255.0))
    rgb, acc, depth, term, wts, _ = ops.composite_forward(*_cuda(raw, mask, z, rays, bg), want_weights=True)
    assert float((wts.sum(1) - acc).abs().max()) < 1e-5                  # acc is the sum of the weights
    assert float(acc.max()) <= 1.0 + 1e-5 and float(wts.min()) >= 0.0    # a partition of unity at most
    assert float(rgb.min()) >= -1e-6 and float(rgb.max()) <= 1.0 + 1e-5  # convex combination with a white background
    assert float((depth - (wts * z.to(d)).sum(1)).abs().max()) < 1e-4
    assert float(acc[1]) == 0.0 and int(term[1]) == 0                    # the empty ray
