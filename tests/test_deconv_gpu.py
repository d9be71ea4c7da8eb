"""GPU: the native transposed-convolution kernels of the motion-weight volume decoder (csrc/deconv.cu) against torch's
conv_transpose3d in fp32 -- the arithmetic of deconv_vol_decoder.py:25-33 / network_util.py:12-50 -- layer by layer (forward, weight,
bias and data gradients, odd channel counts, every volume size of the stack) and for the whole decoder module.
`exact` = 3 x tf32 (what allow_tf32 = False selects): fp32-grade tolerances; plain tf32: the tolerance of a 10-bit mantissa."""
import pytest
import torch
import torch.nn.functional as F

from occnerf_b200 import ops
from tests.helpers import dev, report

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("Cin,Cout,D", [(1024, 512, 1), (96, 40, 2), (70, 25, 4), (64, 64, 8), (256, 25, 16)])
@pytest.mark.parametrize("exact", [True, False])
def test_layer_against_torch(Cin, Cout, D, exact):
    torch.backends.cudnn.allow_tf32 = False
    d = dev()
    gen = torch.Generator().manual_seed(Cin + D)
    W = (torch.randn(Cin, Cout, 4, 4, 4, generator=gen) / (Cin * 8) ** 0.5).to(d)
    b = torch.randn(Cout, generator=gen).to(d)
    Yin = torch.randn(Cin, D ** 3, generator=gen).to(d)
    gY = torch.randn(Cout, (2 * D) ** 3, generator=gen).to(d)
    # reference: fp32 autograd
    Wt, bt, Yt = W.clone().requires_grad_(True), b.clone().requires_grad_(True), Yin.clone().requires_grad_(True)
    out = F.conv_transpose3d(F.leaky_relu(Yt, 0.2).view(1, Cin, D, D, D), Wt, bt, stride=2, padding=1)[0].reshape(Cout, -1)
    (out * gY).sum().backward()
    Yout = ops.deconv3d_forward(W, b, Yin, D, 0.2, exact)
    dW, db, dYin = ops.deconv3d_backward(W, Yin, gY, D, 0.2, exact)
    errs = dict(fwd=_rel(Yout, out.detach()), dW=_rel(dW, Wt.grad), db=_rel(db, bt.grad), dYin=_rel(dYin, Yt.grad))
    report(f"deconv_layer[{Cin}x{Cout}x{D},{'3xtf32' if exact else 'tf32'}]", **errs)
    tol = 2e-5 if exact else 3e-3
    assert max(errs.values()) < tol, errs
    # identity activation (slope 1) and a caller-provided weight-gradient destination
    Yout1 = ops.deconv3d_forward(W, b, Yin, D, 1.0, exact)
    want1 = F.conv_transpose3d(Yin.view(1, Cin, D, D, D), W, b, stride=2, padding=1)[0].reshape(Cout, -1)
    assert _rel(Yout1, want1) < tol
    dst = torch.full_like(W, 7.0)
    dW2, _, none = ops.deconv3d_backward(W, Yin, gY, D, 0.2, exact, need_dyin=False, dW_out=dst)
    assert none is None and dW2.data_ptr() == dst.data_ptr() and _rel(dst, Wt.grad) < tol


@pytest.mark.parametrize("allow_tf32", [False, True])
def test_decoder_module_against_library_path(allow_tf32):
    from occnerf_b200 import prologue as P
    d = dev()
    torch.manual_seed(0)
    dec = P.MotionWeightVolumeDecoder().to(d)
    priors = torch.rand(1, 25, 32, 32, 32, device=d) + 0.01
    gv = torch.randn(1, 25, 32, 32, 32, device=d)
    torch.backends.cudnn.allow_tf32 = False                    # the library cross-check always in fp32
    dec.native = False
    want = dec(motion_weights_priors=priors)
    (want * gv).sum().backward()
    ref = {n: p.grad.clone() for n, p in dec.named_parameters()}
    dec.zero_grad()
    torch.backends.cudnn.allow_tf32 = allow_tf32               # selects tf32 / 3 x tf32 in the native kernels
    dec.native = True
    c0 = ops._lib.COUNTERS["calls"]
    got = dec(motion_weights_priors=priors)
    (got * gv).sum().backward()
    assert ops._lib.COUNTERS["calls"] - c0 >= 14, "the native decoder kernels must have run"
    torch.backends.cudnn.allow_tf32 = True
    e_vol = float((got - want).abs().max())
    e_grads = {n: _rel(p.grad, ref[n]) for n, p in dec.named_parameters()}
    report(f"decoder_module[{'tf32' if allow_tf32 else '3xtf32'}]", vol=e_vol, worst_grad=max(e_grads.values()))
    assert e_vol < (1e-4 if allow_tf32 else 2e-6)
    # (tf32: five layers of 10-bit-mantissa products in a row, and LeakyReLU sign flips of near-zero pre-activations: 2 % measured)
    assert max(e_grads.values()) < (5e-2 if allow_tf32 else 5e-4), e_grads
    # the 56 taps of the first transposed convolution that a 1 x 1 x 1 input never touches receive an exactly zero gradient
    g1 = dec.decoder.block_conv[0].weight.grad
    dead = g1.clone()
    dead[:, :, 1:3, 1:3, 1:3] = 0
    assert float(dead.abs().max()) == 0.0
