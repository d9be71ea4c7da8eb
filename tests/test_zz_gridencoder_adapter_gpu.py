"""GPU: the `_gridencoder` drop-in (occnerf_b200/gridencoder_backend.py) called the way the reference's grid.py calls its
extension (grid.py:48-55 forward: pre-allocated [L,B,C] outputs + optional dy_dx; grid.py:78-85 backward: caller-zeroed
grad_embeddings, optional grad_inputs), against (a) the operator layer used by the render path (same C entry point ->
bitwise equal) and (b) the reference's own gridencoder.cu compiled unmodified into oracle/_ref (same Python signature:
the two modules are called with identical argument lists).  Runs last: it was written after the round's GPU budget was
spent and has not been executed on a B200 yet."""
import pytest
import torch

from occnerf_b200 import _lib, gridencoder_backend as be, ops
from tests.helpers import dev, normwise_close
from tests.test_hashgrid_gpu import _inputs, _load_ref, _table

pytestmark = pytest.mark.gpu


def _call_like_grid_py(mod, x, emb, offs, Sv, g):
    """grid.py:40-55 and :72-85 restated: the exact argument lists the reference passes to its backend."""
    B, D = x.shape
    L, C, H = offs.shape[0] - 1, emb.shape[1], 16
    outputs = torch.empty(L, B, C, device=x.device, dtype=emb.dtype)
    dy_dx = torch.empty(B, L * D * C, device=x.device, dtype=emb.dtype)
    mod.grid_encode_forward(x, emb, offs, outputs, B, D, C, L, Sv, H, dy_dx, 0, False, 0)
    grad_embeddings = torch.zeros_like(emb)
    grad_inputs = torch.zeros_like(x)
    mod.grid_encode_backward(g, x, emb, offs, grad_embeddings, B, D, C, L, Sv, H, dy_dx, grad_inputs, 0, False, 0)
    torch.cuda.synchronize()
    return outputs, dy_dx, grad_embeddings, grad_inputs


def test_adapter_equals_operator_layer():
    emb, offs, Sv = _table(seed=3, scale=0.1)
    d = dev()
    B = 20000
    x, emb_d, offs_d = _inputs(B, seed=4).to(d), emb.to(d), offs.to(d)
    g = torch.randn(16, B, 2, generator=torch.Generator().manual_seed(1)).to(d)
    out, dy_dx, ge, gi = _call_like_grid_py(be, x, emb_d, offs_d, Sv, g)
    sc = ops.level_scales(Sv, 16, 16, d)
    out2, dy2, _, _ = ops.hashgrid_forward(x, emb_d, offs_d, sc, layout=_lib.LAYOUT_LBC, want_dy_dx=True)
    assert torch.equal(out, out2) and torch.equal(dy_dx, dy2)
    assert torch.equal(out.permute(1, 0, 2).reshape(B, 32), ops.hashgrid_forward(x, emb_d, offs_d, sc)[0])   # grid.py:58
    ge2 = torch.zeros_like(emb_d)
    ops.hashgrid_backward(g.data_ptr(), 32, _lib.LAYOUT_LBC, x, offs_d, sc, ge2, 2)
    gi2 = ops.hashgrid_input_backward(g.data_ptr(), 32, _lib.LAYOUT_LBC, dy2, B, 4, 2, 16)
    assert normwise_close(ge.cpu().numpy(), ge2.cpu().numpy(), 1e-5) and torch.equal(ge != 0, ge2 != 0)
    assert torch.equal(gi, gi2)
    # without dy_dx (grid.py:50-53: calc_grad_inputs=False passes None both ways)
    out3 = torch.empty_like(out)
    be.grid_encode_forward(x, emb_d, offs_d, out3, B, 4, 2, 16, Sv, 16, None, 0, False, 0)
    ge3 = torch.zeros_like(emb_d)
    be.grid_encode_backward(g, x, emb_d, offs_d, ge3, B, 4, 2, 16, Sv, 16, None, None, 0, False, 0)
    assert torch.equal(out3, out) and normwise_close(ge3.cpu().numpy(), ge2.cpu().numpy(), 1e-5)


def test_adapter_against_compiled_reference_module():
    ref = _load_ref()
    emb, offs, Sv = _table(seed=6, scale=0.1)
    d = dev()
    B = 30000
    x, emb_d, offs_d = _inputs(B, seed=8).to(d), emb.to(d), offs.to(d)
    g = torch.randn(16, B, 2, generator=torch.Generator().manual_seed(2)).to(d)
    ours = _call_like_grid_py(be, x, emb_d, offs_d, Sv, g)
    theirs = _call_like_grid_py(ref, x, emb_d, offs_d, Sv, g)
    assert int((ours[0].view(torch.int32) != theirs[0].view(torch.int32)).sum()) == 0      # forward bitwise identical
    assert float((ours[1] - theirs[1]).abs().max()) <= 1e-6 * float(theirs[1].abs().max())
    assert normwise_close(ours[2].cpu().numpy(), theirs[2].cpu().numpy(), 1e-5) and torch.equal(ours[2] != 0, theirs[2] != 0)
    assert normwise_close(ours[3].cpu().numpy(), theirs[3].cpu().numpy(), 1e-5)


def test_rays_all_float32_camera_fixture():
    """occnerf_generate_rays in OCCNERF_RAYS_ALL_F32 mode (tpose.py:66-84 cameras) against the fixture the reference's
    functions wrote for a float32 K, R, T.  Lives in this run-last file for the same reason as the tests above: the mode was
    added after the GPU budget ended; its arithmetic is checked bitwise on the host by tests/test_rays_host_emulation_cpu.py.
    Bar: ray_mask identical, o exact, d / near / far within 1e-6 relative."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rays_f32.npz"))
    assert g["K"].dtype == g["R"].dtype == g["T"].dtype == np.float32
    rays, mask, n, _ = ops.generate_rays(int(g["H"]), int(g["W"]), g["K"], g["R"], g["T"], g["bbox_min"], g["bbox_max"])
    rays, mask = rays.cpu().numpy(), mask.cpu().numpy()
    assert np.array_equal(mask, g["ray_mask"]) and n == int(g["ray_mask"].sum()) > 0
    assert np.array_equal(rays[:, 0:3], g["rays_o"])
    for got, want in ((rays[:, 3:6], g["rays_d"]), (rays[:, 6], g["near"]), (rays[:, 7], g["far"])):
        assert np.all(np.abs(got - want) <= 1e-6 * np.maximum(1.0, np.abs(want)))
