"""CPU: the oracle (oracle/occnerf_oracle.py + hashgrid_oracle.c) against the golden fixtures written by
the UNMODIFIED reference (oracle/make_golden.py).  This is what pins the oracle."""
import copy

import numpy as np
import pytest
import torch

from oracle import make_golden, occnerf_oracle as O
from tests.helpers import load_case, normwise_close


def _params_with_grad(sub, w):
    sub, w = copy.deepcopy(sub), copy.deepcopy(w)
    for t in [w.embeddings, sub.point_dist, w.geo_w, w.geo_b, w.out_w, w.out_b] + w.pts_w + w.pts_b + w.rgb_w + w.rgb_b:
        t.requires_grad_(True)
    return sub, w


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_render_matches_reference(name):
    sub, w, fr, vol, t_rand, rk, g = load_case(name)
    sub, w = _params_with_grad(sub, w)
    vol = vol.clone().requires_grad_(True)
    o = O.render_rays(fr, vol, sub, w, iter_val=rk["iter_val"], training=rk["training"], t_rand=t_rand, return_aux=True)
    # stage outputs
    assert np.array_equal(o["z"].detach().numpy(), g["z"])                       # bit-exact sampling
    assert np.abs(o["x_skel"].detach().numpy() - g["x_skel"]).max() < 2e-6
    assert np.abs(o["mask"].detach().numpy() - g["mask"]).max() < 1e-6
    assert np.allclose(o["raw"].detach().numpy(), g["raw"], rtol=1e-5, atol=2e-5)
    assert np.array_equal(o["term"].numpy().astype(np.int32), g["term"])
    # BASELINE.json tolerance is 1e-3 absolute; the restatement is far inside it
    for k in ("rgb", "alpha", "depth"):
        assert np.abs(o[k].detach().numpy() - g[k]).max() < 1e-5, k
    if rk["training"]:
        assert np.abs(o["comp_loss"].detach().numpy() - g["comp_loss"]).max() < 1e-5
        assert np.array_equal(o["hits"].numpy(), g["counter_delta"])
        make_golden.scalar_loss(o).backward()
        ge = w.embeddings.grad.reshape(-1).numpy()
        assert normwise_close(ge[g["g_emb_idx"]], g["g_emb_val"], 1e-3)
        offs = w.offsets.tolist()
        l2 = np.array([w.embeddings.grad[a:b].double().norm().item() for a, b in zip(offs[:-1], offs[1:])])
        assert np.allclose(l2, g["g_emb_level_l2"], rtol=2e-3)
        gv = vol.grad.reshape(-1).numpy()
        assert normwise_close(gv[g["g_vol_idx"]], g["g_vol_val"], 1e-3)
        assert np.isclose(vol.grad.double().norm().item(), float(g["g_vol_l2"]), rtol=1e-4)
        assert normwise_close(sub.point_dist.grad.numpy(), g["g_point_dist"], 1e-3)
        for i in range(4):
            sl = (slice(None, None, 8), slice(None, None, 8)) if i else (slice(None), slice(None))
            for nm, t in ((f"g_pts_w{i}", w.pts_w[i]), (f"g_rgb_w{i}", w.rgb_w[i])):
                assert normwise_close(t.grad.numpy()[sl], g[nm], 1e-3), nm
            for nm, t in ((f"g_pts_b{i}", w.pts_b[i]), (f"g_rgb_b{i}", w.rgb_b[i])):
                assert normwise_close(t.grad.numpy(), g[nm], 1e-3), nm
        for nm, t in (("g_geo_w", w.geo_w), ("g_geo_b", w.geo_b), ("g_out_w", w.out_w), ("g_out_b", w.out_b)):
            assert normwise_close(t.grad.numpy(), g[nm], 1e-3), nm


@pytest.mark.parametrize("name", ["train_dense", "eval_init"])
def test_voxel_bins_bit_exact(name):
    """Integer voxel bins of the 24-bone warp: oracle's explicit-FMA affine + ATen un-normalisation vs the
    reference's own matmul/grid_sample arithmetic."""
    sub, w, fr, vol, t_rand, rk, g = load_case(name)
    z = O.z_samples(fr.near, fr.far, 128, t_rand)
    pts = O.sample_points(fr.rays_o, fr.rays_d, z).reshape(-1, 3)
    _, _, bins = O.lbs_warp(pts, fr.motion_scale_Rs, fr.motion_Ts, vol, fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz,
                            exact=True, return_bins=True)
    assert np.array_equal(bins.numpy().astype(np.int16), g["bins"])


@pytest.mark.parametrize("name", ["zju", "f64", "f32"])
def test_rays_oracle_matches_reference(name):
    """oracle/rays_oracle.py against get_rays_from_KRT + rays_intersect_3d_bbox of the reference (camera_util.py:133-212),
    fixtures by oracle/make_golden_rays.py.  float64 intermediates are compared exactly: same numpy calls, same dtypes."""
    import os
    from oracle import rays_oracle as RO
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"rays_{name}.npz"))
    H, W = int(g["H"]), int(g["W"])
    o, d = RO.pixel_rays(H, W, g["K"], g["R"], g["T"])
    o, d = np.ascontiguousarray(o.reshape(-1, 3)), np.ascontiguousarray(d.reshape(-1, 3))
    near, far, mask = RO.box_near_far(g["bbox_min"], g["bbox_max"], o, d)
    assert np.array_equal(mask, g["ray_mask"])                                   # integer decision: bit-exact
    assert 0 < mask.sum() < mask.size
    assert np.abs(d[mask] - g["rays_d_f64"]).max() <= 1e-15
    assert np.abs(near - g["near_f64"]).max() <= 1e-14 and np.abs(far - g["far_f64"]).max() <= 1e-14
    packed, mask2, pix = RO.frame_rays(H, W, g["K"], g["R"], g["T"], g["bbox_min"], g["bbox_max"])
    assert packed.dtype == np.float32 and np.array_equal(mask2, mask) and np.array_equal(pix, np.nonzero(mask)[0])
    assert np.array_equal(packed[:, 0:3], g["rays_o"]) and np.array_equal(packed[:, 3:6], g["rays_d"])
    assert np.array_equal(packed[:, 6], g["near"]) and np.array_equal(packed[:, 7], g["far"])


def test_image_oracle_matches_reference():
    """oracle/image_oracle.py against run.py:39-66 unpack_to_image (+ image_util.to_8b_image), fixture by
    oracle/make_golden_rays.py: 8-bit frames are integer work -> bit-exact."""
    import os
    from oracle import image_oracle as IO
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "image_unpack.npz"))
    rgb8, alpha8 = IO.unpack(int(g["W"]), int(g["H"]), g["ray_mask"], g["bgcolor"] / 255., g["rgb"], g["alpha"])
    assert np.array_equal(rgb8, g["rgb_image"])
    assert np.array_equal(alpha8, g["alpha_image"][..., 0]) and np.array_equal(alpha8, g["alpha_image"][..., 2])
    assert 0 in rgb8 and 255 in rgb8
