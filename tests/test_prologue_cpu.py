"""CPU: the per-frame prologue that `Network.forward` runs in front of the ray path (network.py:551-597; inside `e2e`,
outside `value`) -- occnerf_b200/prologue.py is library (torch) code, so it runs here -- against the outputs of the
UNMODIFIED reference modules (MotionBasisComputer network_util.py:138-200, MotionWeightVolumeDecoder
deconv_vol_decoder.py:25-33 + network_util.py:12-50, BodyPoseRefiner mlp_delta_body_pose.py:35-41) carrying the same,
seeded parameters (tests/golden/prologue.npz by oracle/make_golden_prologue.py).  Tolerance 1e-5 absolute on rotations /
translations (fp32, different but equivalent matrix-chain order) and 1e-6 on the softmax-ed weight volume."""
import os

import numpy as np
import torch

from oracle import make_golden_prologue as G

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prologue.npz")


def test_prologue_matches_reference_modules():
    g = np.load(GOLDEN)
    sub, fr = G.inputs()
    pro = G.seeded_prologue()
    assert np.array_equal(fr.dst_Rs.numpy(), g["dst_Rs"]) and np.array_equal(fr.dst_posevec.numpy(), g["dst_posevec"])
    assert np.isclose(sum(float(p.detach().double().sum()) for p in pro.parameters()), float(g["param_checksum"]), rtol=1e-9)
    with torch.no_grad():
        for tag, iter_val in (("plain", 500), ("refined", 10 ** 7)):                      # kick-in at 1000 in the seeded module
            Rs, Ts, vol = pro(fr.dst_Rs, fr.dst_Ts, fr.cnl_gtfms, sub.priors, fr.dst_posevec, iter_val)
            assert np.abs(Rs.numpy() - g[f"motion_scale_Rs_{tag}"]).max() < 1e-5, tag
            assert np.abs(Ts.numpy() - g[f"motion_Ts_{tag}"]).max() < 1e-5, tag
        refined = pro.pose_decoder(fr.dst_posevec[None])["Rs"]
    assert np.abs(refined.numpy() - g["refined_Rs"]).max() < 1e-6
    assert np.abs(g["motion_scale_Rs_refined"] - g["motion_scale_Rs_plain"]).max() > 1e-3       # the refinement branch did something
    assert list(vol.shape) == g["vol_shape"].tolist()
    assert np.abs(vol[:, ::4, ::4, ::4].numpy() - g["vol_sub"]).max() < 1e-6
    assert np.abs(vol.double().sum(dim=(1, 2, 3)).numpy() - g["vol_channel_sums"]).max() < 1e-3 * np.abs(g["vol_channel_sums"]).max()


def test_network_prologue_hook_equals_the_module():
    """`Network.install_prologue()` wires the same three modules under the reference's attribute names."""
    from occnerf_b200 import synthetic as S
    from occnerf_b200.network import RenderConfig
    sub, fr = G.inputs()
    pro = G.seeded_prologue()
    net = S.network_from_synthetic(sub, S.make_weights(sub.bound, seed=0), RenderConfig(), device="cpu")
    net.install_prologue(pose_kick_in_iter=1000)
    for name in ("motion_basis_computer", "mweight_vol_decoder", "pose_decoder"):
        getattr(net, name).load_state_dict(getattr(pro, name).state_dict(), strict=True)
    with torch.no_grad():
        for iter_val in (500, 10 ** 7):
            a = pro(fr.dst_Rs, fr.dst_Ts, fr.cnl_gtfms, sub.priors, fr.dst_posevec, iter_val)
            b = net._run_prologue(fr.dst_Rs, fr.dst_Ts, fr.cnl_gtfms, sub.priors, fr.dst_posevec, iter_val)
            assert all(torch.equal(x, y) for x, y in zip(a, b))
