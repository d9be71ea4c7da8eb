"""GPU: hash-grid kernels (occnerf_hashgrid_*) against the C oracle (bit-exact cells/slots), and against the
reference's own CUDA kernels compiled unmodified into oracle/_ref (when that build is present)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from occnerf_b200 import _lib, ops, synthetic as S
from oracle import hashgrid_c
from tests.helpers import dev, maxabs, normwise_close, report

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_gridencoder_ref.so")


def _table(seed=0, scale=0.05, bound=1.2):
    offs, pls = S.hashgrid_offsets(desired_resolution=2048 * bound)
    gen = torch.Generator().manual_seed(seed)
    emb = (torch.rand(int(offs[-1]), 2, generator=gen) * 2 - 1) * scale
    return emb, torch.from_numpy(offs), float(np.log2(pls))


def _inputs(B, seed=1, oob=True):
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 4, generator=gen)
    if B >= 8:
        x[0] = 0.0
        x[1] = 1.0
        x[2] = torch.tensor([0.0, 1.0, 0.5, 0.25])
        if oob:
            x[3, 1] = 1.0000001
            x[4, 0] = -1e-7
            x[5, 3] = 7.0
    return x


def test_cells_and_slots_bit_exact_and_values():
    emb, offs, Sv = _table()
    d = dev()
    scales = ops.level_scales(Sv, 16, 16, d)
    host_scales = hashgrid_c.host_level_scales(Sv, 16, 16)
    ulps = (scales.cpu().view(torch.int32) - host_scales.view(torch.int32)).abs().max().item()
    x = _inputs(20000)
    out, dy_dx, cells, slots = ops.hashgrid_forward(x.to(d), emb.to(d), offs.to(d), scales, want_dy_dx=True, want_cells=True)
    r = hashgrid_c.forward(x, emb, offs, Sv, 16, want_dy_dx=True, want_cells=True, level_scales=scales.cpu().contiguous())
    assert torch.equal(cells.cpu(), r["cells"]), "integer cell coordinates differ"
    assert torch.equal(slots.cpu(), r["idx"]), "hash-table slots differ"
    eo, ed = maxabs(out, r["out"]), maxabs(dy_dx, r["dy_dx"]) / float(r["dy_dx"].abs().max())
    report("hashgrid_fwd", out=eo, dy_dx_rel=ed, scale_ulps_vs_host=ulps)
    assert eo < 1e-7 and ed < 1e-5
    assert float(out[3:6].abs().max()) == 0.0, "out-of-range samples encode to zero (gridencoder.cu:110-135)"


def test_backward_and_input_backward():
    emb, offs, Sv = _table(scale=0.1)
    d = dev()
    scales = ops.level_scales(Sv, 16, 16, d)
    x = _inputs(6000, seed=3)
    g = torch.randn(6000, 32, generator=torch.Generator().manual_seed(4))
    g_emb = torch.zeros_like(emb).to(d)
    g_d, x_d = g.to(d), x.to(d)
    ops.hashgrid_backward(g_d.data_ptr(), 32, _lib.LAYOUT_BLC, x_d, offs.to(d), scales, g_emb, 2)
    ref32, ref64 = hashgrid_c.backward(g, x, offs, emb.shape[0], 2, Sv, 16, level_scales=scales.cpu().contiguous(), want_f64=True)
    e = maxabs(g_emb, ref64) / float(ref64.abs().max())
    report("hashgrid_bwd", g_emb_rel_vs_f64=e, oracle_f32_rel_vs_f64=maxabs(ref32, ref64) / float(ref64.abs().max()))
    assert e < 1e-5
    assert torch.equal((g_emb != 0).cpu(), ref64 != 0) or float((g_emb.cpu() - ref64.float()).abs().max()) < 1e-6
    # run-length variants (consecutive samples sharing a cell are merged before the reduction): same sums.
    # Inputs that really share cells: short random walks, plus out-of-range rows and zero gradients in the middle of runs
    gen = torch.Generator().manual_seed(11)
    xw = (torch.rand(400, 1, 4, generator=gen) + torch.cumsum(torch.randn(400, 16, 4, generator=gen) * 2e-3, 1)).reshape(-1, 4)
    xw[37] = 1.5
    xw[1000:1003] = -0.1
    gw = torch.randn(6400, 32, generator=gen)
    gw[5::7] = 0.0
    # ragged tail: a sample count that is not a multiple of the run length
    xw, gw = xw[:6397].contiguous(), gw[:6397].contiguous()
    refw32, refw64 = hashgrid_c.backward(gw, xw, offs, emb.shape[0], 2, Sv, 16, level_scales=scales.cpu().contiguous(), want_f64=True)
    gw_d, xw_d = gw.to(d), xw.to(d).contiguous()
    # the run-length FORWARD kernel returns bitwise what the per-sample kernel returns (strided output, out-of-range rows)
    e_d, o_d = emb.to(d), offs.to(d)
    f_plain = torch.full((6397, 40), -3.0, device=d)
    f_runs = torch.full((6397, 40), -3.0, device=d)
    ops.hashgrid_forward(xw_d, e_d, o_d, scales, out_ptr=f_plain.data_ptr() + 16, ld=40)
    ops.hashgrid_forward(xw_d, e_d, o_d, scales, out_ptr=f_runs.data_ptr() + 16, ld=40, run_length=16)
    assert torch.equal(f_plain, f_runs) and float(f_runs[:, :4].max()) == -3.0 and float(f_runs[37, 4:36].abs().max()) == 0.0
    for rl in (0, 8, 16):
        ge = torch.zeros_like(emb).to(d)
        ops.hashgrid_backward(gw_d.data_ptr(), 32, _lib.LAYOUT_BLC, xw_d, offs.to(d), scales, ge, 2, run_length=rl)
        e = maxabs(ge, refw64) / float(refw64.abs().max())
        report(f"hashgrid_bwd_runs{rl}", g_emb_rel_vs_f64=e)
        assert e < 1e-5, (rl, e)
    # input gradient
    _o, dy_dx, _c, _s = ops.hashgrid_forward(x.to(d), emb.to(d), offs.to(d), scales, want_dy_dx=True)
    gi = ops.hashgrid_input_backward(g_d.data_ptr(), 32, _lib.LAYOUT_BLC, dy_dx, 6000, 4, 2, 16)
    gi_o = hashgrid_c.input_backward(g, dy_dx.cpu(), 6000, 4, 2, 16)
    assert normwise_close(gi.cpu().numpy(), gi_o.numpy(), 1e-5)


def test_autograd_function_and_strided_output():
    emb, offs, Sv = _table()
    d = dev()
    x = _inputs(777, seed=9, oob=False).to(d).requires_grad_(True)
    e = emb.to(d).requires_grad_(True)
    out = ops.grid_encode(x, e, offs.to(d), float(2.0 ** Sv), 16)
    g = torch.randn(777, 32, generator=torch.Generator().manual_seed(1)).to(d)
    out.backward(g)
    xo, eo = x.detach().cpu().requires_grad_(True), emb.clone().requires_grad_(True)
    scales = ops.level_scales(Sv, 16, 16, d).cpu().contiguous()
    oo = hashgrid_c.HashGridFn.apply(xo, eo, offs, Sv, 16, scales)
    oo.backward(g.cpu())
    assert maxabs(out, oo) < 1e-7
    assert normwise_close(e.grad.cpu().numpy(), eo.grad.numpy(), 1e-5)
    assert normwise_close(x.grad.cpu().numpy(), xo.grad.numpy(), 1e-4)
    # encode straight into a wider row (the MLP input buffer): ld = 132, column offset 100
    XB = torch.full((777, 132), -1.0, device=d)
    sc = ops.level_scales(Sv, 16, 16, d)
    ops.hashgrid_forward(x.detach(), emb.to(d), offs.to(d), sc, out_ptr=XB.data_ptr() + 400, ld=132)
    assert torch.equal(XB[:, 100:], out.detach()) and float(XB[:, :100].max()) == -1.0


@pytest.mark.parametrize("D,C", [(2, 2), (3, 2), (3, 4), (4, 1), (4, 8), (2, 1)])
def test_other_shapes_against_oracle(D, C):
    offs, pls = S.hashgrid_offsets(input_dim=D, num_levels=8, log2_hashmap_size=14, desired_resolution=512)
    gen = torch.Generator().manual_seed(D * 10 + C)
    emb = torch.rand(int(offs[-1]), C, generator=gen) - 0.5
    x = torch.rand(3000, D, generator=gen)
    d = dev()
    Sv = float(np.log2(pls))
    sc = ops.level_scales(Sv, 16, 8, d)
    offs_t = torch.from_numpy(offs)
    out, dy_dx, cells, slots = ops.hashgrid_forward(x.to(d), emb.to(d), offs_t.to(d), sc, want_dy_dx=True, want_cells=True)
    r = hashgrid_c.forward(x, emb, offs_t, Sv, 16, want_dy_dx=True, want_cells=True, level_scales=sc.cpu().contiguous())
    assert torch.equal(slots.cpu(), r["idx"]) and torch.equal(cells.cpu(), r["cells"])
    assert maxabs(out, r["out"]) < 1e-6
    g = torch.randn(3000, 8 * C, generator=gen)
    g_emb = torch.zeros_like(emb).to(d)
    g_d, x_d = g.to(d), x.to(d)
    ops.hashgrid_backward(g_d.data_ptr(), 8 * C, 0, x_d, offs_t.to(d), sc, g_emb, C)
    ref32, _ = hashgrid_c.backward(g, x, offs_t, emb.shape[0], C, Sv, 16, level_scales=sc.cpu().contiguous())
    assert normwise_close(g_emb.cpu().numpy(), ref32.numpy(), 1e-5)


def test_unsupported_and_empty():
    emb, offs, Sv = _table()
    d = dev()
    sc = ops.level_scales(Sv, 16, 16, d)
    out, *_ = ops.hashgrid_forward(torch.zeros(0, 4, device=d), emb.to(d), offs.to(d), sc)
    assert out.shape == (0, 32)
    with pytest.raises(RuntimeError, match="unsupported D"):
        ops.hashgrid_forward(torch.zeros(4, 5, device=d), emb.to(d), offs.to(d), sc)


def _load_ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/_gridencoder_ref.so not built (run oracle/build_ref.sh where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("_gridencoder_ref", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_against_compiled_reference_kernels():
    """The reference's own gridencoder.cu, compiled unmodified for sm_100a, on the same inputs: outputs must be
    bitwise identical (same arithmetic, same slots); the atomically accumulated gradient agrees to fp32 ordering noise.
    This is what pins the hash-grid restatement (oracle + CUDA) to the real reference."""
    ref = _load_ref()
    emb, offs, Sv = _table(seed=5, scale=0.1)
    d = dev()
    B = 50000
    x = _inputs(B, seed=11).to(d)
    emb_d, offs_d = emb.to(d), offs.to(d)
    out_ref = torch.empty(16, B, 2, device=d)
    dy_ref = torch.empty(B, 16 * 4 * 2, device=d)
    ref.grid_encode_forward(x, emb_d, offs_d, out_ref, B, 4, 2, 16, Sv, 16, dy_ref, 0, False, 0)
    torch.cuda.synchronize()
    sc = ops.level_scales(Sv, 16, 16, d)
    out_lbc, dy_dx, _, _ = ops.hashgrid_forward(x, emb_d, offs_d, sc, layout=_lib.LAYOUT_LBC, want_dy_dx=True)
    out_blc, *_ = ops.hashgrid_forward(x, emb_d, offs_d, sc)
    nbit = int((out_lbc.view(torch.int32) != out_ref.view(torch.int32)).sum())
    report("hashgrid_vs_refkernel", differing_words=nbit, maxabs=maxabs(out_lbc, out_ref), dy_dx=maxabs(dy_dx, dy_ref))
    assert nbit == 0, "forward output is not bitwise identical to the reference kernel"
    assert torch.equal(out_blc, out_ref.permute(1, 0, 2).reshape(B, 32))
    assert maxabs(dy_dx, dy_ref) <= 1e-6 * float(dy_ref.abs().max())
    g = torch.randn(16, B, 2, device=d)
    ge_ref = torch.zeros_like(emb_d)
    gi_ref = torch.zeros(B, 4, device=d)
    ref.grid_encode_backward(g, x, emb_d, offs_d, ge_ref, B, 4, 2, 16, Sv, 16, dy_ref, gi_ref, 0, False, 0)
    torch.cuda.synchronize()
    ge = torch.zeros_like(emb_d)
    ops.hashgrid_backward(g.data_ptr(), 32, _lib.LAYOUT_LBC, x, offs_d, sc, ge, 2)
    gi = ops.hashgrid_input_backward(g.data_ptr(), 32, _lib.LAYOUT_LBC, dy_dx, B, 4, 2, 16)
    assert normwise_close(ge.cpu().numpy(), ge_ref.cpu().numpy(), 1e-5)
    assert torch.equal(ge != 0, ge_ref != 0)
    assert normwise_close(gi.cpu().numpy(), gi_ref.cpu().numpy(), 1e-5)


def test_linearity_and_adjoint_at_full_size():
    """BASELINE config 2 size (786 432 samples): encode is linear in the table, and backward is its adjoint."""
    emb, offs, Sv = _table(seed=2, scale=1.0)
    d = dev()
    B = 6144 * 128
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(B, 4, generator=gen).to(d)
    sc = ops.level_scales(Sv, 16, 16, d)
    e1, offs_d = emb.to(d), offs.to(d)
    e2 = torch.randn(emb.shape, generator=gen).to(d)
    f1, *_ = ops.hashgrid_forward(x, e1, offs_d, sc)
    f2, *_ = ops.hashgrid_forward(x, e2, offs_d, sc)
    f12, *_ = ops.hashgrid_forward(x, (2.0 * e1 - 0.5 * e2).contiguous(), offs_d, sc)
    assert float((f12 - (2.0 * f1 - 0.5 * f2)).abs().max()) < 1e-5
    g = torch.randn(B, 32, generator=gen).to(d)
    ge = torch.zeros_like(e1)
    ops.hashgrid_backward(g.data_ptr(), 32, 0, x, offs_d, sc, ge, 2)
    lhs, rhs = float((f2.double() * g.double()).sum()), float((e2.double() * ge.double()).sum())
    report("hashgrid_adjoint_786k", lhs=lhs, rhs=rhs)
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0)
