"""GPU: the fused per-vertex block (occnerf_vertex_block_*, network.py:263-284 + occnerf_mlp.py:171-175) against the
oracle's restatement and against torch autograd of the same formula (gradients to point_dist and to the hash table)."""
import pytest
import torch

from occnerf_b200 import ops, synthetic as S
from occnerf_b200.network import RenderConfig
from oracle import occnerf_oracle as O
from tests.helpers import dev, maxabs, normwise_close, report

pytestmark = pytest.mark.gpu


def _torch_vertex_features(net):
    """The block as the reference writes it, in eager torch on the device (autograd provides the reference gradients)."""
    st = net._static()
    V = net.point_base.shape[0]
    pc = net.point_base + net.point_dist
    kidx = ops.knn(pc.detach().contiguous(), st["base4"], [0, V], 3)[:, 0].long()
    b = net.point_base[kidx]
    direction = pc[:, None, :] - b
    n = net.point_norms[kidx]
    a = torch.abs(torch.nn.functional.cosine_similarity(direction, n, dim=-1))[..., None]
    knn_base = (a * b).sum(1) / a.sum(1)
    inside = ((direction * n).sum(-1) < 0).sum(1) > 1.5
    dist = direction.norm(dim=-1).mean(1, keepdim=True)
    dist = torch.where(inside[:, None], -dist, dist)
    v_in = torch.cat([(knn_base + net.bound) / (2 * net.bound), torch.clamp((dist + 0.2) / 0.8, 0.0, 1.0)], -1)
    hv = net.cnl_mlp.module.encoder(v_in, bound=None)      # already in [0, 1]
    return torch.cat([hv, pc, torch.zeros(V, 1, device=pc.device, dtype=pc.dtype)], -1), pc, knn_base, dist, v_in, kidx


def test_vertex_block_forward_and_gradients():
    d = dev()
    sub = S.make_subject(seed=0)
    w = S.make_weights(sub.bound, seed=0, table_scale=0.05)
    net = S.network_from_synthetic(sub, w, RenderConfig(), device=d)
    with torch.no_grad():                       # move the cloud off the base so that the geometry is not degenerate
        net.point_dist.copy_((torch.rand(net.point_dist.shape, generator=torch.Generator().manual_seed(2)) * 2 - 1).to(d) * 3e-3)
    ref, pc_ref, kb_ref, dist_ref, v_in_ref, kidx = _torch_vertex_features(net)
    pc_o, kb_o, dist_o = O.vertex_block(net.point_base.detach().cpu(), net.point_dist.detach().cpu(), net.point_norms.cpu())
    assert maxabs(kb_ref, kb_o) < 1e-5 and maxabs(dist_ref, dist_o) < 1e-6     # the torch formula IS the oracle's
    st = net._static()
    v_in = torch.empty(6890, 4, device=d)
    tail = torch.empty(6890, 4, device=d)
    ops.vertex_block_forward(st["point_base"], net.point_dist.detach().reshape(-1).contiguous(), st["point_norms"],
                             kidx.to(torch.int32).contiguous(), net.bound, v_in, tail.data_ptr(), 4)
    e_v = maxabs(v_in, v_in_ref)
    assert e_v < 5e-7, e_v                       # the hash-grid input itself: fp32 rounding only
    got, pc = net.vertex_features()
    e_f, e_pc = maxabs(got, ref), maxabs(pc, pc_ref)
    # (the finest hash level has ~3700 cells per unit: a 1e-7 difference in v_in moves the features by ~1e-5 x table scale)
    assert e_pc == 0.0 and e_f < 1e-4, (e_f, e_pc)
    assert float(got[:, 35].abs().max()) == 0.0
    # the geometry backward kernel in isolation: given d loss / d v_in and d loss / d pc, torch autograd of the formula
    # must give the same d loss / d point_dist (no hash grid in between)
    gen = torch.Generator().manual_seed(5)
    gv, gt = torch.randn(6890, 4, generator=gen).to(d), torch.randn(6890, 4, generator=gen).to(d)
    (v_in_ref * gv).sum().backward(retain_graph=True, inputs=[net.point_dist])
    g1 = net.point_dist.grad.clone()
    net.point_dist.grad = None
    (pc_ref * gt[:, :3]).sum().backward(retain_graph=True, inputs=[net.point_dist])
    g_ref = (g1 + net.point_dist.grad).reshape(-1)
    net.point_dist.grad = None
    g_got = ops.vertex_block_backward(st["point_base"], net.point_dist.detach().reshape(-1).contiguous(), st["point_norms"],
                                      kidx.to(torch.int32).contiguous(), net.bound, gv.contiguous(), gt.data_ptr(), 4)
    e_k = maxabs(g_got, g_ref) / float(g_ref.abs().max())
    assert e_k < 2e-4, e_k          # (|pc - b| ~ 5e-3 is a difference of O(1) coordinates: ~2e-5 relative precision in fp32)
    # end to end through the hash grid (features and their input gradients are ill-conditioned at the finest levels: a
    # 1e-7 difference in v_in crosses a cell boundary for a few of the 6890 vertices, so compare in the L2 norm)
    g = torch.randn(got.shape, generator=gen).to(d)
    emb = net.cnl_mlp.module.encoder.embeddings
    (ref * g).sum().backward()
    g_pd_ref, g_emb_ref = net.point_dist.grad.clone(), emb.grad.clone()
    net.point_dist.grad, emb.grad = None, None
    (got * g).sum().backward()
    e_pd = float((net.point_dist.grad - g_pd_ref).norm() / g_pd_ref.norm())
    e_emb = float((emb.grad - g_emb_ref).norm() / g_emb_ref.norm())
    report("vertex_block", feats=e_f, v_in=e_v, g_kernel_rel=e_k, g_point_dist_l2=e_pd, g_emb_l2=e_emb)
    assert e_pd < 5e-2 and e_emb < 1e-2, (e_pd, e_emb)
