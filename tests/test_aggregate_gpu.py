"""GPU: visibility-weighted aggregation (occnerf_aggregate_*) against the oracle; both backward kernels."""
import pytest
import torch

from occnerf_b200 import ops, synthetic as S
from oracle import occnerf_oracle as O
from tests.helpers import dev, load_case, maxabs, normwise_close, report

pytestmark = pytest.mark.gpu


def _case(n_rays=96, S_=128):
    sub, w, fr, vol, t_rand, rk, g = load_case("train_dense")
    xyz = torch.from_numpy(g["x_skel"]).reshape(-1, 3)[: n_rays * S_].contiguous()
    idx = O.multiscale_knn(xyz, sub.point_base, sub.fps_index, 10, stable=False)
    gen = torch.Generator().manual_seed(0)
    feats = torch.randn(6890, 35, generator=gen)
    counter = torch.ones(6890) + (torch.rand(6890, generator=gen) < 0.3).float() * torch.randint(0, 9, (6890,), generator=gen)
    return idx, feats, counter


def test_forward_against_oracle():
    idx, feats, counter = _case()
    m = idx.shape[0]
    att, var = O.visibility_attention(counter, idx)
    want = (att[..., None] * feats[idx.reshape(m, -1)]).sum(1)
    d = dev()
    f36 = torch.zeros(6890, 36, device=d)
    f36[:, :35] = feats.to(d)
    X = torch.full((m, 132), -3.0, device=d)
    ops.aggregate_forward(idx.to(torch.int32).to(d).contiguous(), counter.to(d), f36, X.data_ptr() + 4 * 64, 132)
    e1, e2 = maxabs(X[:, 64:99], want), maxabs(X[:, 99], var[:, 0])
    report("aggregate_fwd", agg=e1, var=e2)
    assert e1 < 5e-6 and e2 < 1e-6
    assert float(X[:, :64].max()) == -3.0 and float(X[:, 100:].max()) == -3.0


@pytest.mark.parametrize("copies", [1, 7, 64])
def test_backward_against_oracle(copies):
    idx, feats, counter = _case()
    m = idx.shape[0]
    fr = feats.clone().requires_grad_(True)
    att, _ = O.visibility_attention(counter, idx)
    agg = (att.detach()[..., None] * fr[idx.reshape(m, -1)]).sum(1)
    g = torch.randn(m, 35, generator=torch.Generator().manual_seed(1))
    (agg * g).sum().backward()
    d = dev()
    gX = torch.randn(m, 132, device=d)                     # columns outside 64..98 must be ignored (99 = variance slot)
    gX[:, 64:99] = g.to(d)
    gf = ops.aggregate_backward(idx.to(torch.int32).to(d).contiguous(), counter.to(d), gX.data_ptr() + 4 * 64, 132, 6890,
                                copies=copies)
    e = maxabs(gf[:, :35], fr.grad) / float(fr.grad.abs().max())
    report(f"aggregate_bwd[copies={copies}]", rel=e)
    assert e < 1e-5
    assert float(gf[:, 35].abs().max()) == 0.0


def test_run_length_backward_matches_per_sample_kernel():
    """Run-length backward (one reduction per run of equal vertices in a neighbour slot, attention weights handed over by
    the forward pass) against the per-sample kernel.  Ragged tail (m not a multiple of 16) included; the forward output
    does not depend on whether the weights are exported."""
    idx, feats, counter = _case()
    m = idx.shape[0] - 5
    idx = idx[:m]
    d = dev()
    f36 = torch.zeros(6890, 36, device=d)
    f36[:, :35] = feats.to(d)
    idx_d, cnt_d = idx.to(torch.int32).to(d).contiguous(), counter.to(d)
    Xa = torch.full((m, 132), -3.0, device=d)
    Xb = torch.full((m, 132), -3.0, device=d)
    ops.aggregate_forward(idx_d, cnt_d, f36, Xa.data_ptr() + 4 * 64, 132)
    att_w = ops.aggregate_forward(idx_d, cnt_d, f36, Xb.data_ptr() + 4 * 64, 132, want_att=True)
    assert torch.equal(Xa, Xb)
    assert float((att_w.sum(1) - 1).abs().max()) < 1e-5
    same_slot = float((idx[1:] == idx[:-1]).float().mean())
    assert same_slot > 0.3, "the case must actually contain runs"
    gX = torch.randn(m, 132, device=d)
    gX[::5, 64:100] = 0.0
    ref = ops.aggregate_backward(idx_d, cnt_d, gX.data_ptr() + 4 * 64, 132, 6890, copies=4)
    got = ops.aggregate_backward(idx_d, cnt_d, gX.data_ptr() + 4 * 64, 132, 6890, copies=4, att_w=att_w)
    e = maxabs(got, ref) / float(ref.abs().max())
    report("aggregate_runs", bwd_rel=e, same_slot=same_slot)
    assert e < 1e-5 and float(got[:, 35].abs().max()) == 0.0
