"""Stress sweep (BASELINE.json configs[4]): the HBM-bound kernels of the path at 2^16..2^20 rays x 64/128/256 samples,
reported as algorithmic GB/s (SURVEY.md 8d bytes per sample) against the measured copy bandwidth.
Writes a markdown table to stdout (committed under profiles/)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from occnerf_b200 import ops, synthetic as S
from occnerf_b200.network import RenderConfig
d = torch.device("cuda")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
HBM = peaks["hbm_gbs"]
sub = S.make_subject(0)
net = S.network_from_synthetic(sub, S.make_weights(sub.bound), RenderConfig(), device=d)
vol = S.make_motion_weights_vol(sub.priors, 0).to(d)
fr0 = S.frame_to(S.make_frame(sub, mode="patch", n_patches=6, patch=32, seed=100), d)
st = net._static()
enc = net.cnl_mlp.module.encoder
scales = ops.level_scales(float(np.log2(enc.per_level_scale)), enc.base_resolution, enc.num_levels, d)
flush = torch.empty(64 * 1024 * 1024, device=d)

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1.0)                      # 256 MiB write: evicts L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n

rows = []
max_samples = int(os.environ.get("MAX_SAMPLES", 2 ** 27))
for logn in (16, 18):
    for Sn in (64, 128, 256):
        N = 2 ** logn
        M = N * Sn
        if M > max_samples:
            continue
        # rays drawn over the view frustum of the synthetic camera: tile the 6144 patch rays with jitter
        idx = torch.randint(0, fr0.rays_o.shape[0], (N,), device=d)
        jit = 1.0 + 0.01 * torch.randn(N, 3, device=d)
        rays = torch.cat([fr0.rays_o[idx], fr0.rays_d[idx] * jit, fr0.near[idx], fr0.far[idx]], -1).contiguous()
        t_rand = torch.rand(N, Sn, device=d)
        Rs, Ts = fr0.motion_scale_Rs.contiguous(), fr0.motion_Ts.contiguous()
        res = {}
        res["warp_fwd"] = (t(lambda: ops.warp_forward(rays, t_rand, Rs, Ts, vol, fr0.cnl_bbox_min_xyz, fr0.cnl_bbox_scale_xyz, Sn)), 20.0 + 32.0 / Sn + 4.0)
        z, x, mask = ops.warp_forward(rays, t_rand, Rs, Ts, vol, fr0.cnl_bbox_min_xyz, fr0.cnl_bbox_scale_xyz, Sn)
        g_mask = torch.randn(N, Sn, device=d)
        res["warp_bwd"] = (t(lambda: ops.warp_backward(rays, t_rand, Rs, Ts, fr0.cnl_bbox_min_xyz, fr0.cnl_bbox_scale_xyz, g_mask, Sn, tuple(vol.shape))), 8.0)
        raw = torch.randn(N, Sn, 5, device=d)
        bg = torch.zeros(3, device=d)
        res["composite_fwd"] = (t(lambda: ops.composite_forward(raw, mask, z, rays, bg, want_comp=True)), 28.0 + 4.0 + 28.0 / Sn)
        g_rgb, g_acc, g_depth, g_comp = torch.randn(N, 3, device=d), torch.randn(N, device=d), torch.randn(N, device=d), torch.randn(N, Sn, device=d)
        res["composite_bwd"] = (t(lambda: ops.composite_backward(raw, mask, z, rays, bg, g_rgb, g_acc, g_depth, g_comp)), 52.0)
        # the real encoder input of these samples: multi-scale KNN -> surface projection + signed distance (smooth along a ray)
        xyz = x.reshape(-1, 3).contiguous()
        grid = ops.build_knn_grid(st["point_base"], [f.to(d) for f in net.fps_index])
        res["knn_grid(4 levels x k=10)"] = (t(lambda: ops.knn_grid(xyz, Sn, grid), n=2), 12.0 + 160.0)
        knn_idx = ops.knn_grid(xyz, Sn, grid)
        res["sample_geometry"] = (t(lambda: ops.sample_geometry(xyz, knn_idx, st["point_base"], st["point_norms"], net.bound), n=2), 12.0 + 40.0 + 20.0)
        enc_in, _ = ops.sample_geometry(xyz, knn_idx, st["point_base"], st["point_norms"], net.bound)
        del knn_idx
        out = torch.empty(M, 32, device=d)
        res["hashgrid_fwd"] = (t(lambda: ops.hashgrid_forward(enc_in, enc.embeddings.detach(), enc.offsets, scales, out=out)), 144.0)
        g = torch.randn(M, 32, device=d)
        g_emb = torch.zeros_like(enc.embeddings)
        res["hashgrid_fwd_run16"] = (t(lambda: ops.hashgrid_forward(enc_in, enc.embeddings.detach(), enc.offsets, scales, out=out, run_length=16)), 144.0)
        res["hashgrid_bwd_run16"] = (t(lambda: ops.hashgrid_backward(g.data_ptr(), 32, 0, enc_in, enc.offsets, scales, g_emb, 2, run_length=16)), 144.0)
        for k, (ms, bps) in res.items():
            gbs = bps * M / (ms * 1e-3) / 1e9
            rows.append((logn, Sn, k, ms, bps, gbs, gbs / HBM))
        del rays, t_rand, z, x, mask, raw, enc_in, out, g, g_mask
        torch.cuda.empty_cache()
print(f"Peak used: measured copy bandwidth {HBM:.1f} GB/s (MEASURED_PEAKS.json).  L2 flushed (256 MiB write) before every timed launch.\n")
print("| rays | samples/ray | kernel | ms | algorithmic B/sample | GB/s | fraction of HBM peak |")
print("|---:|---:|---|---:|---:|---:|---:|")
for logn, Sn, k, ms, bps, gbs, frac in rows:
    print(f"| 2^{logn} | {Sn} | {k} | {ms:.3f} | {bps:.2f} | {gbs:.0f} | {frac:.3f} |")
