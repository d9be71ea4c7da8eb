"""GPU: image assembly behind the path (csrc/image.cu, occnerf_unpack_image through the C ABI) against the reference's
unpack_to_image (run.py:39-66, fixture written by the reference's own function text) and the numpy oracle -- 8-bit frames
are integer work: bit-exact --, then `render.render_view` end to end: rays generated on the device, rendered through
Network.forward, painted on the device; three "ranks" run one after the other must merge into the single-rank frame."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_unpack_against_reference_fixture():
    from occnerf_b200 import ops
    g = np.load(os.path.join(GOLDEN, "image_unpack.npz"))
    H, W = int(g["H"]), int(g["W"])
    d = torch.device("cuda:0")
    pix = torch.from_numpy(np.nonzero(g["ray_mask"])[0].astype(np.int32)).to(d)
    rgb8, alpha8, bad = ops.unpack_image(torch.from_numpy(g["rgb"]).to(d), torch.from_numpy(g["alpha"]).to(d), pix, H, W, g["bgcolor"] / 255.)
    assert int(bad.item()) == 0
    assert np.array_equal(rgb8.cpu().numpy(), g["rgb_image"])
    assert np.array_equal(alpha8.cpu().numpy(), g["alpha_image"][..., 0])


def test_unpack_against_oracle_full_size_shards_and_errors():
    from occnerf_b200 import ops
    from oracle import image_oracle as IO
    H = W = 1024
    rng = np.random.default_rng(0)
    mask = rng.uniform(size=H * W) < 0.7
    n = int(mask.sum())
    rgb = rng.uniform(-0.5, 1.5, (n, 3)).astype(np.float32)
    alpha = rng.uniform(-0.5, 1.5, n).astype(np.float32)
    bg = np.array([10.0, 200.0, 128.0]) / 255.
    want_rgb, want_alpha = IO.unpack(W, H, mask, bg, rgb, alpha)
    d = torch.device("cuda:0")
    pix = torch.from_numpy(np.nonzero(mask)[0].astype(np.int32)).to(d)
    t_rgb, t_alpha = torch.from_numpy(rgb).to(d), torch.from_numpy(alpha).to(d)
    rgb8, alpha8, bad = ops.unpack_image(t_rgb, t_alpha, pix, H, W, bg)
    assert int(bad.item()) == 0 and np.array_equal(rgb8.cpu().numpy(), want_rgb) and np.array_equal(alpha8.cpu().numpy(), want_alpha)
    # two shards into one frame: fill + scatter, then scatter only
    h = n // 3
    f_rgb, f_alpha, _ = ops.unpack_image(t_rgb[:h].contiguous(), t_alpha[:h].contiguous(), pix[:h].contiguous(), H, W, bg)
    ops.unpack_image(t_rgb[h:].contiguous(), t_alpha[h:].contiguous(), pix[h:].contiguous(), H, W, bg, out=(f_rgb, f_alpha), fill=False)
    assert torch.equal(f_rgb, rgb8) and torch.equal(f_alpha, alpha8)
    # no rays at all: the background frame
    e_rgb, e_alpha, _ = ops.unpack_image(t_rgb[:0], t_alpha[:0], pix[:0], 8, 8, bg)
    assert np.array_equal(e_rgb.cpu().numpy(), IO.unpack(8, 8, np.zeros(64, bool), bg, rgb[:0], alpha[:0])[0]) and not e_alpha.any().item()
    # pixel indices outside the frame are counted, not written
    _, _, bad = ops.unpack_image(t_rgb[:4].contiguous(), t_alpha[:4].contiguous(), torch.tensor([0, 64, -1, 5], dtype=torch.int32, device=d), 8, 8, bg)
    assert int(bad.item()) == 2
    with pytest.raises(RuntimeError, match="bad sizes"):
        ops.unpack_image(t_rgb[:4].contiguous(), None, pix[:4].contiguous(), 0, 8, bg)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.unpack_image(torch.zeros(4, 3), None, pix[:4].contiguous(), 8, 8, bg)


def test_render_view_sharded_equals_single():
    from occnerf_b200 import render, synthetic as S
    from occnerf_b200.network import RenderConfig
    from oracle import image_oracle as IO
    d = torch.device("cuda:0")
    sub = S.make_subject(seed=0)
    w = S.make_weights(sub.bound, seed=0, table_scale=0.05, nonzero_bias=True)
    net = S.network_from_synthetic(sub, w, RenderConfig(perturb=0.0, mlp_engine="fp32"), device=d).train(False)
    net.install_prologue()
    fr = S.frame_to(S.make_frame(sub, mode="patch", n_patches=1, patch=8, seed=9), d)
    H = W = 48
    K, R, T = S.lookat_camera(W)
    R, T = R.astype(np.float64), T.astype(np.float64)
    box = {"min_xyz": np.array([-0.95, -1.35, -0.45], np.float32), "max_xyz": np.array([0.95, 0.65, 0.45], np.float32)}
    data = dict(dst_Rs=fr.dst_Rs, dst_Ts=fr.dst_Ts, cnl_gtfms=fr.cnl_gtfms, motion_weights_priors=sub.priors.to(d), dst_posevec=fr.dst_posevec,
                cnl_bbox_min_xyz=fr.cnl_bbox_min_xyz, cnl_bbox_scale_xyz=fr.cnl_bbox_scale_xyz, bgcolor=fr.bgcolor)
    bg = [0.0, 0.0, 0.0]
    one = render.render_view(net, H, W, K, R, T, box, data, bg)
    b, e, n = one["rays"]
    assert (b, e) == (0, n) and n > 100 and int(one["bad_pixels"].item()) == 0
    mask = one["ray_mask"].cpu().numpy()
    assert int(mask.sum()) == n
    rgb8 = one["rgb8"].cpu().numpy()
    assert rgb8.shape == (H, W, 3) and rgb8.dtype == np.uint8
    assert not rgb8.reshape(-1, 3)[~mask].any()                       # background outside the box silhouette
    frames, spans = [], []
    for r in range(3):
        part = render.render_view(net, H, W, K, R, T, box, data, bg, rank=r, world=3)
        frames.append((part["rgb8"].cpu().numpy(), part["alpha8"].cpu().numpy()))
        spans.append(part["rays"][:2])
    assert spans[0][0] == 0 and spans[-1][1] == n
    m_rgb, m_alpha = render.merge_frames(frames, mask, spans)
    # rays are independent; the 8-bit frames may differ by one step where a value sits on a rounding edge
    assert np.abs(m_rgb.astype(int) - rgb8.astype(int)).max() <= 1
    assert np.abs(m_alpha.astype(int) - one["alpha8"].cpu().numpy().astype(int)).max() <= 1
