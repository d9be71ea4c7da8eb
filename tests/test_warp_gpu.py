"""GPU: fused sample+warp kernel (occnerf_warp_*) against the golden fixtures and the oracle."""
import numpy as np
import pytest
import torch

from occnerf_b200 import ops, synthetic as S
from oracle import make_golden, occnerf_oracle as O
from tests.helpers import dev, load_case, maxabs, normwise_close, report

pytestmark = pytest.mark.gpu


def _run(fr, vol, t_rand, want_bins=True):
    d = dev()
    rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).to(d).contiguous()
    return ops.warp_forward(rays, t_rand.to(d).contiguous() if t_rand is not None else None, fr.motion_scale_Rs.to(d).contiguous(),
                            fr.motion_Ts.to(d).contiguous(), vol.to(d).contiguous(), fr.cnl_bbox_min_xyz.to(d),
                            fr.cnl_bbox_scale_xyz.to(d), 128, want_bins=want_bins), rays


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_against_reference_golden(name):
    sub, w, fr, vol, t_rand, rk, g = load_case(name)
    (z, x_skel, mask, bins), _ = _run(fr, vol, t_rand)
    assert np.array_equal(z.cpu().numpy(), g["z"]), "sample depths must be bit-exact"
    assert np.array_equal(bins.cpu().numpy().reshape(g["bins"].shape).astype(np.int16), g["bins"]), "voxel bins must be bit-exact"
    ex, em = maxabs(x_skel, g["x_skel"]), maxabs(mask, g["mask"])
    report(f"warp_golden[{name}]", x_skel=ex, mask=em)
    assert ex < 2e-6 and em < 1e-6


def test_against_oracle_large_and_backward():
    sub = S.make_subject(seed=0)
    fr = S.make_frame(sub, mode="patch", n_patches=6, patch=32, seed=21)
    vol = S.make_motion_weights_vol(sub.priors, seed=2)
    N = fr.rays_o.shape[0]
    t_rand = torch.rand(N, 128, generator=torch.Generator().manual_seed(5))
    (z, x_skel, mask, bins), rays = _run(fr, vol, t_rand)
    volr = vol.clone().requires_grad_(True)
    zo = O.z_samples(fr.near, fr.far, 128, t_rand)
    pts = O.sample_points(fr.rays_o, fr.rays_d, zo).reshape(-1, 3)
    xo, mo, bo = O.lbs_warp(pts, fr.motion_scale_Rs, fr.motion_Ts, volr, fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz,
                            exact=True, return_bins=True)
    assert torch.equal(z.cpu(), zo)
    assert torch.equal(bins.cpu().reshape(-1, 24, 3), bo)
    ex, em = maxabs(x_skel.reshape(-1, 3), xo), maxabs(mask.reshape(-1), mo)
    gm = torch.randn(N, 128, generator=torch.Generator().manual_seed(6))
    (mo * gm.reshape(-1)).sum().backward()
    d = dev()
    g_vol = ops.warp_backward(rays, t_rand.to(d).contiguous(), fr.motion_scale_Rs.to(d).contiguous(), fr.motion_Ts.to(d).contiguous(),
                              fr.cnl_bbox_min_xyz.to(d), fr.cnl_bbox_scale_xyz.to(d), gm.to(d).contiguous(), 128, tuple(vol.shape))
    eg = maxabs(g_vol, volr.grad) / float(volr.grad.abs().max())
    report("warp_oracle_6144", x_skel=ex, mask=em, g_vol_rel=eg)
    assert ex < 2e-6 and em < 1e-6
    assert normwise_close(g_vol.cpu().numpy(), volr.grad.numpy(), 1e-4)
    assert float(g_vol[24].abs().max()) == 0.0, "the background channel never receives a gradient (network.py:363)"


def test_no_jitter_and_odd_sizes():
    sub = S.make_subject(seed=0)
    fr = S.make_frame(sub, mode="image", img=64, max_rays=37, seed=4)
    vol = S.make_motion_weights_vol(sub.priors, seed=1)
    d = dev()
    rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).to(d).contiguous()
    for Sn in (64, 100, 256):
        z, x_skel, mask = ops.warp_forward(rays, None, fr.motion_scale_Rs.to(d).contiguous(), fr.motion_Ts.to(d).contiguous(),
                                           vol.to(d).contiguous(), fr.cnl_bbox_min_xyz.to(d), fr.cnl_bbox_scale_xyz.to(d), Sn)
        zo = O.z_samples(fr.near, fr.far, Sn, None)
        xo, mo = O.lbs_warp(O.sample_points(fr.rays_o, fr.rays_d, zo).reshape(-1, 3), fr.motion_scale_Rs, fr.motion_Ts, vol,
                            fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz)
        assert torch.equal(z.cpu(), zo)
        assert maxabs(x_skel.reshape(-1, 3), xo) < 2e-6 and maxabs(mask.reshape(-1), mo) < 1e-6
    # empty batch
    z, x_skel, mask = ops.warp_forward(rays[:0].contiguous(), None, fr.motion_scale_Rs.to(d).contiguous(), fr.motion_Ts.to(d).contiguous(),
                                       vol.to(d).contiguous(), fr.cnl_bbox_min_xyz.to(d), fr.cnl_bbox_scale_xyz.to(d), 128)
    assert z.shape == (0, 128)


def test_packed_kernels_equal_the_scalar_kernels_bitwise():
    """The corner-packed path (vol8 + TMA-staged block inputs, what Network uses) against the scalar-gather kernel on the
    reference layout: the same rounding sequence, so z / x_skel / mask are bit-identical -- with and without jitter, for S that
    takes the bulk-copy path (multiple of 4) and S that does not."""
    sub = S.make_subject(seed=0)
    vol = S.make_motion_weights_vol(sub.priors, seed=2)
    d = dev()
    for mode, kw, Sn in (("patch", dict(n_patches=6, patch=32), 128), ("image", dict(img=64, max_rays=301), 33), ("image", dict(img=64, max_rays=77), 130)):
        fr = S.make_frame(sub, mode=mode, seed=21, **kw)
        N = fr.rays_o.shape[0]
        rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).to(d).contiguous()
        args = (fr.motion_scale_Rs.to(d).contiguous(), fr.motion_Ts.to(d).contiguous(), vol.to(d).contiguous(), fr.cnl_bbox_min_xyz.to(d),
                fr.cnl_bbox_scale_xyz.to(d), Sn)
        for t_rand in (None, torch.rand(N, Sn, generator=torch.Generator().manual_seed(5)).to(d)):
            a = ops.warp_forward(rays, t_rand, *args)
            b = ops.warp_forward(rays, t_rand, *args, want_bins=True)
            for x, y, nm in zip(a, b[:3], ("z", "x_skel", "mask")):
                assert torch.equal(x, y), (nm, mode, Sn, t_rand is not None)
    # backward: packed (vector reductions + fold) vs scalar reductions
    fr = S.make_frame(sub, mode="patch", n_patches=6, patch=32, seed=21)
    N = fr.rays_o.shape[0]
    rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).to(d).contiguous()
    t_rand = torch.rand(N, 128, generator=torch.Generator().manual_seed(5)).to(d)
    gm = torch.randn(N, 128, generator=torch.Generator().manual_seed(6)).to(d)
    bw = (rays, t_rand, fr.motion_scale_Rs.to(d).contiguous(), fr.motion_Ts.to(d).contiguous(), fr.cnl_bbox_min_xyz.to(d),
          fr.cnl_bbox_scale_xyz.to(d), gm, 128, tuple(vol.shape))
    g1, g0 = ops.warp_backward(*bw), ops.warp_backward(*bw, packed=False)
    assert normwise_close(g1.cpu().numpy(), g0.cpu().numpy(), 1e-5)
    assert float(g1[24].abs().max()) == 0.0


def test_pose_gradients_against_autograd():
    """d mask / d (motion_scale_Rs, motion_Ts): what F.grid_sample's grid gradient gives the reference (network.py:367-370)."""
    sub = S.make_subject(seed=0)
    fr = S.make_frame(sub, mode="patch", n_patches=2, patch=16, seed=8)
    vol = S.make_motion_weights_vol(sub.priors, seed=2)
    N = fr.rays_o.shape[0]
    t_rand = torch.rand(N, 128, generator=torch.Generator().manual_seed(5))
    Rs, Ts = fr.motion_scale_Rs.clone().requires_grad_(True), fr.motion_Ts.clone().requires_grad_(True)
    volr = vol.clone().requires_grad_(True)
    zo = O.z_samples(fr.near, fr.far, 128, t_rand)
    pts = O.sample_points(fr.rays_o, fr.rays_d, zo).reshape(-1, 3)
    _, mo = O.lbs_warp(pts, Rs, Ts, volr, fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz, exact=False)
    gm = torch.randn(N, 128, generator=torch.Generator().manual_seed(6))
    (mo * gm.reshape(-1)).sum().backward()
    d = dev()
    rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).to(d).contiguous()
    vol8 = ops.warp_pack_volume(vol.to(d).contiguous(), 24)
    g_vol, g_Rs, g_Ts = ops.warp_backward(rays, t_rand.to(d), fr.motion_scale_Rs.to(d).contiguous(), fr.motion_Ts.to(d).contiguous(),
                                          fr.cnl_bbox_min_xyz.to(d), fr.cnl_bbox_scale_xyz.to(d), gm.to(d), 128, tuple(vol.shape), vol8=vol8,
                                          want_pose=True)
    eR = maxabs(g_Rs, Rs.grad) / float(Rs.grad.abs().max())
    eT = maxabs(g_Ts, Ts.grad) / float(Ts.grad.abs().max())
    report("warp_pose_grads", g_Rs_rel=eR, g_Ts_rel=eT, scale_R=float(Rs.grad.abs().max()))
    assert eR < 2e-4 and eT < 2e-4
    assert normwise_close(g_vol.cpu().numpy(), volr.grad.numpy(), 1e-4)
    # and through the autograd node of the network path
    from occnerf_b200.network import _WarpFn
    Rd, Td = fr.motion_scale_Rs.to(d).clone().requires_grad_(True), fr.motion_Ts.to(d).clone().requires_grad_(True)
    vd_ = vol.to(d).clone().requires_grad_(True)
    z, xs, mask = _WarpFn.apply(vd_, rays, t_rand.to(d), Rd, Td, fr.cnl_bbox_min_xyz.to(d), fr.cnl_bbox_scale_xyz.to(d), 128, None)
    (mask * gm.to(d)).sum().backward()
    assert normwise_close(Rd.grad.cpu().numpy(), Rs.grad.numpy(), 2e-4) and normwise_close(Td.grad.cpu().numpy(), Ts.grad.numpy(), 2e-4)
