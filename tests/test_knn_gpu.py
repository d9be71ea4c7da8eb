"""GPU: exact KNN + per-sample surface geometry (occnerf_knn, occnerf_sample_geometry) against the oracle."""
import numpy as np
import pytest
import torch

from occnerf_b200 import ops, synthetic as S
from oracle import occnerf_oracle as O
from tests.helpers import dev, load_case, maxabs, report

pytestmark = pytest.mark.gpu


def _subject_arrays(sub):
    d = dev()
    base = sub.point_base.to(d)
    fps = [f.to(d) for f in sub.fps_index]
    sup = torch.cat([base] + [base[f] for f in fps], 0)
    gid = torch.cat([torch.arange(base.shape[0], device=d)] + fps).to(torch.int32).contiguous()
    lb = np.cumsum([0, base.shape[0]] + [int(f.shape[0]) for f in fps]).tolist()
    return ops.to_float4(sup), gid, lb


def test_multiscale_ids_exact_on_golden_points():
    sub, w, fr, vol, t_rand, rk, g = load_case("train_dense")
    xyz = torch.from_numpy(g["x_skel"]).reshape(-1, 3)[:6000].contiguous()
    sup4, gid, lb = _subject_arrays(sub)
    idx = ops.knn(xyz.to(dev()), sup4, lb, 10, support_gid=gid)
    ref = O.multiscale_knn(xyz, sub.point_base, sub.fps_index, 10)
    assert idx.shape == (6000, 4, 10)
    assert torch.equal(idx.cpu().long(), ref), "neighbour ids differ from the exact brute-force oracle"


def test_random_queries_ties_k3_and_selection():
    gen = torch.Generator().manual_seed(0)
    sup = torch.rand(5000, 3, generator=gen)
    sup[100:110] = sup[50:60]              # exact duplicates -> distance ties, lower row must win
    q = torch.cat([torch.rand(3000, 3, generator=gen) * 1.4 - 0.2, sup[:200]], 0).contiguous()
    d = dev()
    sup4 = ops.to_float4(sup.to(d))
    for k in (3, 10):
        idx = ops.knn(q.to(d), sup4, [0, 5000], k)[:, 0]
        ref = O.knn_bruteforce(q, sup, k)
        assert torch.equal(idx.cpu().long(), ref)
    # a support point queried against its own set finds itself (or its lower-index duplicate) first
    self_idx = ops.knn(sup.to(d).contiguous(), sup4, [0, 5000], 3)[:, 0, 0].cpu().long()
    expect = torch.arange(5000)
    expect[100:110] = torch.arange(50, 60)
    assert torch.equal(self_idx, expect)
    # query selection leaves unselected rows untouched
    sel = (torch.arange(q.shape[0]) % 3 == 0).to(torch.uint8)
    out = torch.full((q.shape[0], 1, 10), -7, dtype=torch.int32, device=d)
    ops.knn(q.to(d), sup4, [0, 5000], 10, query_sel=sel.to(d), out=out)
    ref = O.knn_bruteforce(q, sup, 10)
    assert torch.equal(out[:, 0].cpu().long()[sel.bool()], ref[sel.bool()])
    assert bool((out[:, 0].cpu()[~sel.bool()] == -7).all())


def test_sample_geometry_against_oracle():
    sub, w, fr, vol, t_rand, rk, g = load_case("train_dense")
    xyz = torch.from_numpy(g["x_skel"]).reshape(-1, 3)[:8000].contiguous()
    sup4, gid, lb = _subject_arrays(sub)
    d = dev()
    idx = ops.knn(xyz.to(d), sup4, lb, 10, support_gid=gid)
    raw = torch.zeros(8000, 5, device=d)
    enc_in, dist = ops.sample_geometry(xyz.to(d), idx, sub.point_base.to(d), sub.point_norms.to(d), sub.bound, raw=raw)
    enc_o, dist_o = O.sample_geometry(xyz, O.knn_bruteforce(xyz, sub.point_base, 10), sub.point_base, sub.point_norms, sub.bound)
    e1, e2 = maxabs(enc_in, enc_o), maxabs(raw[:, 4], dist_o[:, 0])
    report("sample_geometry", enc_in=e1, dist=e2)
    assert e1 < 1e-6 and e2 < 1e-6
    assert float(raw[:, :4].abs().max()) == 0.0
    assert float(enc_in.min()) >= 0.0 and float(enc_in.max()) <= 1.0


def test_sortedness_at_full_size():
    """BASELINE config 2 size: 786 432 queries; distances to the returned neighbours are ascending on every level."""
    sub = S.make_subject(seed=0)
    sup4, gid, lb = _subject_arrays(sub)
    d = dev()
    gen = torch.Generator().manual_seed(1)
    q = (torch.rand(6144 * 128, 3, generator=gen) * 2.4 - 1.2).to(d)
    idx = ops.knn(q, sup4, lb, 10, support_gid=gid).long()
    base = sub.point_base.to(d)
    dist = (q[:, None, None, :] - base[idx]).norm(dim=-1)
    assert bool((dist[..., 1:] >= dist[..., :-1] - 1e-6).all())
    assert int(idx.min()) >= 0 and int(idx.max()) < base.shape[0]
    lvl3 = set(sub.fps_index[2].tolist())
    assert set(idx[:1000, 3].reshape(-1).tolist()) <= lvl3


def test_hierarchical_search_is_bit_identical_to_brute_force():
    """occnerf_knn_hier (cluster-pruned) must return exactly the ids of the brute-force kernel / oracle on all 4 levels,
    for points on the body, far outside it, and with duplicated support points (ties)."""
    sub = S.make_subject(seed=0)
    d = dev()
    base = sub.point_base.to(d).clone()
    fps = [f.to(d) for f in sub.fps_index]
    sup4, gid, lb = _subject_arrays(sub)
    gen = torch.Generator().manual_seed(3)
    sub_, w, fr, vol, t_rand, rk, g = load_case("train_dense")
    near = torch.from_numpy(g["x_skel"]).reshape(-1, 3)
    far = torch.rand(20000, 3, generator=gen) * 6 - 3
    exact = sub.point_base[torch.randint(0, 6890, (3000,), generator=gen)]
    zeros = torch.zeros(500, 3)
    for S_ in (128, 1, 7):
        q = torch.cat([near, far, exact, zeros], 0).contiguous().to(d)
        ref = ops.knn(q, sup4, lb, 10, support_gid=gid)
        out = torch.full((q.shape[0], 4, 10), -5, dtype=torch.int32, device=d)
        h0 = ops.build_knn_hierarchy(base, base[fps[1]])
        h1 = ops.build_knn_hierarchy(base[fps[0]], base[fps[2]])
        ops.knn_hier(q, S_, *h0, out, 0, 2, None, fps[1].to(torch.int32).contiguous())
        ops.knn_hier(q, S_, *h1, out, 1, 3, fps[0].to(torch.int32).contiguous(), fps[2].to(torch.int32).contiguous())
        bad = (out != ref).any(-1)
        assert not bool(bad.any()), (S_, int(bad.sum()), bad.nonzero()[:5].tolist())
        tree = ops.build_knn_tree(base, fps)
        for lane_rays in (32, 8, 1):           # the warp -> query mapping is a scheduling hint only
            out_t = ops.knn_tree(q, S_, tree, lane_rays=lane_rays)
            bad = (out_t != ref).any(-1)
            assert not bool(bad.any()), ("tree", S_, lane_rays, int(bad.sum()), bad.nonzero()[:5].tolist())
        grid = ops.build_knn_grid(base, fps)
        out_g = ops.knn_grid(q, S_, grid)            # `far` and part of `near` lie outside the grid: exhaustive fallback
        bad = (out_g != ref).any(-1)
        assert not bool(bad.any()), ("grid", S_, int(bad.sum()), bad.nonzero()[:5].tolist())
    ref_cpu = O.multiscale_knn(near[:4000], sub.point_base, sub.fps_index, 10)
    assert torch.equal(out[:4000].cpu().long(), ref_cpu)
