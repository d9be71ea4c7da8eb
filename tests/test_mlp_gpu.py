"""GPU: canonical MLP engines.  The exact-fp32 SIMT path against torch (CPU oracle), and the fused tcgen05/TMEM
kernel against the fp32 path, forward and (hybrid) backward."""
import numpy as np
import pytest
import torch

from occnerf_b200 import mlp as M, synthetic as S
from oracle import occnerf_oracle as O
from tests.helpers import dev, maxabs, normwise_close, report

pytestmark = pytest.mark.gpu


def _weights(seed=0, bias=True):
    sub = S.make_subject(seed=0)
    w = S.make_weights(sub.bound, seed=seed, nonzero_bias=bias)
    return w


def _flat(w, d):
    t = []
    for i in range(4):
        t += [w.pts_w[i], w.pts_b[i]]
    t += [w.geo_w, w.geo_b]
    for i in range(4):
        t += [w.rgb_w[i], w.rgb_b[i]]
    t += [w.out_w, w.out_b]
    return M.MlpWeights([x.to(d) for x in t])


def _inputs(m, seed=0):
    gen = torch.Generator().manual_seed(seed)
    agg = torch.randn(m, 35, generator=gen) * 0.5
    var = torch.rand(m, 1, generator=gen) * 0.1
    h = torch.randn(m, 32, generator=gen) * 0.05
    return agg, var, h


def _oracle(agg, var, h, w):
    rgb, sigma = O.canonical_mlp(agg, var, h, w)
    return torch.cat([rgb, sigma], -1)


def _engine(name):
    if name == "fp32":
        return M.MlpSimt()
    from occnerf_b200 import mlp_tc
    return mlp_tc.MlpTc(n_pass={"tc1": 1, "tf32": 2, "tc3": 3}[name])


@pytest.mark.parametrize("engine,tol", [("fp32", 2e-5), ("tc3", 1e-4), ("tf32", 5e-3), ("tc1", 3e-1)])
@pytest.mark.parametrize("m", [128, 1000, 20000])
def test_forward_against_torch(engine, tol, m):
    w = _weights()
    agg, var, h = _inputs(m, seed=m)
    want = _oracle(agg, var, h, w)
    d = dev()
    XB = torch.zeros(m, 132, device=d)
    XB[:, 64:99], XB[:, 99:100], XB[:, 100:] = agg.to(d), var.to(d), h.to(d)
    raw = torch.full((m, 5), 7.0, device=d)
    _engine(engine).forward(XB, raw, _flat(w, d), save=False)
    e = maxabs(raw[:, :4], want)
    report(f"mlp_fwd[{engine},{m}]", err=e, scale=float(want.abs().max()))
    assert e < tol * max(1.0, float(want.abs().max()))
    assert float(raw[:, 4].min()) == 7.0 and float(raw[:, 4].max()) == 7.0      # the dist channel is not touched


@pytest.mark.parametrize("engine,rel", [("fp32", 1e-4), ("tc3", 1e-2), ("tf32", 8e-2), ("tc1", 3e-1)])
def test_backward_against_torch(engine, rel):
    m = 3000
    w = _weights(seed=2)
    agg, var, h = _inputs(m, seed=5)
    leaves = [agg, h] + w.pts_w + w.pts_b + [w.geo_w, w.geo_b] + w.rgb_w + w.rgb_b + [w.out_w, w.out_b]
    for t in leaves:
        t.requires_grad_(True)
    out = _oracle(agg, var, h, w)
    g = torch.randn(m, 4, generator=torch.Generator().manual_seed(1))
    (out * g).sum().backward()
    d = dev()
    XB = torch.zeros(m, 132, device=d)
    XB[:, 64:99], XB[:, 99:100], XB[:, 100:] = agg.detach().to(d), var.to(d), h.detach().to(d)
    raw = torch.zeros(m, 5, device=d)
    eng = _engine(engine)
    W = _flat(w, d)
    saved = eng.forward(XB, raw, W, save=True)
    g_raw = torch.zeros(m, 5, device=d)
    g_raw[:, :4] = g.to(d)
    gXB, grads = eng.backward(XB, g_raw, W, saved)
    worst = 0.0
    names = M.MlpWeights.ORDER
    ref = {}
    for i in range(4):
        ref[f"pts_w{i}"], ref[f"pts_b{i}"], ref[f"rgb_w{i}"], ref[f"rgb_b{i}"] = w.pts_w[i].grad, w.pts_b[i].grad, w.rgb_w[i].grad, w.rgb_b[i].grad
    ref["geo_w"], ref["geo_b"], ref["out_w"], ref["out_b"] = w.geo_w.grad, w.geo_b.grad, w.out_w.grad, w.out_b.grad
    # fp32: max-norm.  tc3: the forward agrees to ~2e-5, which flips the ReLU mask of the handful of units whose
    # pre-activation is that close to zero; one flipped unit moves one row of a weight gradient by O(1/sqrt(m)) of
    # its maximum, so the tensor-core engine is compared in the Frobenius norm (any re-ordered fp32 GEMM has this).
    def err(a, b):
        a, b = a.detach().cpu().double(), b.detach().cpu().double()
        return float((a - b).abs().max() / b.abs().max()) if engine == "fp32" else float((a - b).norm() / b.norm())
    for nme, gt in zip(names, grads):
        e = err(gt, ref[nme])
        worst = max(worst, e)
        assert e < rel, (nme, e)
    e_agg = err(gXB[:, 64:99], agg.grad)
    e_h = err(gXB[:, 100:132], h.grad)
    report(f"mlp_bwd[{engine}]", worst_param=worst, g_agg=e_agg, g_h=e_h)
    assert e_agg < rel and e_h < rel


def test_wgrad_kernel_matches_library_gemm():
    """occnerf_mlp_wgrad_tc (TMA + MN-major tcgen05) against cuBLAS on the very same bf16 operands."""
    from occnerf_b200 import mlp_tc
    m = 5000                                   # not a multiple of 64: exercises the zero-padded tail
    w = _weights(seed=4)
    agg, var, h = _inputs(m, seed=9)
    d = dev()
    XB = torch.zeros(m, 132, device=d)
    XB[:, 64:99], XB[:, 99:100], XB[:, 100:] = agg.to(d), var.to(d), h.to(d)
    W = _flat(w, d)
    g_raw = torch.zeros(m, 5, device=d)
    g_raw[:, :4] = torch.randn(m, 4, generator=torch.Generator().manual_seed(3)).to(d)
    out = {}
    for mode in ("tc", "lib"):
        eng = mlp_tc.MlpTc(3, wgrad=mode)
        raw = torch.zeros(m, 5, device=d)
        XBc = XB.clone()                        # the forward writes the geometry features into columns 0..63
        saved = eng.forward(XBc, raw, W, save=True)
        out[mode] = eng.backward(XBc, g_raw, W, saved)
    assert torch.equal(out["tc"][0][:, 64:], out["lib"][0][:, 64:])      # columns 0..63 of gXB are not produced
    worst = 0.0
    for nme, a, b in zip(M.MlpWeights.ORDER, out["tc"][1], out["lib"][1]):
        e = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        worst = max(worst, e)
        assert a.shape == b.shape and e < 5e-3, (nme, e)      # cuBLAS returns bf16-rounded products
    report("mlp_wgrad_tc_vs_cublas", worst=worst)
