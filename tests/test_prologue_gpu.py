"""GPU: the native per-frame prologue stages (csrc/prologue.cu: motion basis, pose refiner + Rodrigues, weight-volume softmax)
against the outputs of the UNMODIFIED reference modules (tests/golden/prologue.npz, written by oracle/make_golden_prologue.py
from MotionBasisComputer network_util.py:138-200, BodyPoseRefiner mlp_delta_body_pose.py:35-41, MotionWeightVolumeDecoder
deconv_vol_decoder.py:25-33).  The decoder's transposed convolutions are library calls (cuDNN) on this path; the fixture pins
the whole module output all the same."""
import os

import numpy as np
import pytest
import torch

from occnerf_b200 import ops
from oracle import make_golden_prologue as G
from tests.helpers import dev, maxabs, report

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prologue.npz")


def test_native_prologue_matches_reference_modules():
    g = np.load(GOLDEN)
    sub, fr = G.inputs()
    pro = G.seeded_prologue().to(dev())
    torch.backends.cudnn.allow_tf32 = False              # compare the library convolutions in fp32
    d = dev()
    with torch.no_grad():
        for tag, iter_val in (("plain", 500), ("refined", 10 ** 7)):
            _lib_calls = dict(ops._lib.COUNTERS)
            Rs, Ts, vol = pro(fr.dst_Rs.to(d), fr.dst_Ts.to(d), fr.cnl_gtfms.to(d), sub.priors.to(d), fr.dst_posevec.to(d), iter_val)
            assert ops._lib.COUNTERS["calls"] - _lib_calls["calls"] >= (3 if tag == "refined" else 2), "the native stages must have run"
            eR, eT = maxabs(Rs, g[f"motion_scale_Rs_{tag}"]), maxabs(Ts, g[f"motion_Ts_{tag}"])
            report(f"prologue_native[{tag}]", Rs=eR, Ts=eT)
            assert eR < 1e-5 and eT < 1e-5, tag
    ev = maxabs(vol[:, ::4, ::4, ::4], g["vol_sub"])
    report("prologue_native[volume]", vol=ev)
    assert ev < 2e-6
    assert np.abs(vol.double().sum(dim=(1, 2, 3)).cpu().numpy() - g["vol_channel_sums"]).max() < 1e-3 * np.abs(g["vol_channel_sums"]).max()


def test_stage_kernels_against_torch():
    d = dev()
    gen = torch.Generator().manual_seed(0)
    # weight-volume softmax and its gradient (with exact zeros in the prior: log -> -inf -> weight 0)
    logits = torch.randn(25, 32, 32, 32, generator=gen)
    priors = torch.rand(25, 32, 32, 32, generator=gen)
    priors[3, :8] = 0.0
    lt = logits.clone().requires_grad_(True)
    want = torch.softmax(lt + torch.log(priors), dim=0)
    gv = torch.randn(25, 32, 32, 32, generator=gen)
    (want * gv).sum().backward()
    vol = ops.weight_volume_forward(logits.to(d), priors.to(d))
    gl = ops.weight_volume_backward(vol, gv.to(d))
    assert maxabs(vol, want) < 1e-6 and float(vol[3, :8].abs().max()) == 0.0
    assert maxabs(gl, lt.grad) < 1e-6
    # motion basis against the torch formulation of the same module (library path of occnerf_b200.prologue)
    from occnerf_b200 import prologue as P, synthetic as S
    sub = S.make_subject(seed=0)
    for seed in range(3):
        fr = S.make_frame(sub, mode="patch", n_patches=1, patch=8, seed=seed, pose_std=0.6)
        Rw, Tw = P.MotionBasisComputer()(fr.dst_Rs[None], fr.dst_Ts[None], fr.cnl_gtfms[None])       # CPU: torch form
        Rs, Ts = ops.motion_basis(fr.dst_Rs.to(d).contiguous(), fr.dst_Ts.to(d).contiguous(), fr.cnl_gtfms.to(d).contiguous())
        assert maxabs(Rs, Rw[0]) < 2e-6 and maxabs(Ts, Tw[0]) < 2e-6
        assert maxabs(Rs, fr.motion_scale_Rs) < 1e-5 and maxabs(Ts, fr.motion_Ts) < 1e-5              # the generator's own (numpy) chain


def test_volume_softmax_gradient_reaches_the_decoder():
    """Through the autograd node: d loss / d const_embedding equals the all-torch formulation."""
    from occnerf_b200 import prologue as P
    torch.backends.cudnn.allow_tf32 = False
    d = dev()
    torch.manual_seed(0)
    dec = P.MotionWeightVolumeDecoder().to(d)
    priors = torch.rand(1, 25, 32, 32, 32, device=d) + 0.01
    gv = torch.randn(1, 25, 32, 32, 32, device=d)
    (dec(motion_weights_priors=priors) * gv).sum().backward()
    g_native = dec.const_embedding.grad.clone()
    dec.zero_grad()
    want = torch.softmax(dec.decoder(dec.const_embedding[None]) + torch.log(priors), dim=1)
    (want * gv).sum().backward()
    assert float((g_native - dec.const_embedding.grad).abs().max()) <= 5e-4 * float(dec.const_embedding.grad.abs().max())   # (native 3 x tf32 decoder against fp32 torch)
