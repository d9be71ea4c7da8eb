"""Round-1c experiments on the real bench workload: KNN variants and hash-grid backward per level."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import ops, synthetic as S, mlp as M
from occnerf_b200.network import RenderConfig
d = torch.device("cuda")
sub = S.make_subject(0)
net = S.network_from_synthetic(sub, S.make_weights(sub.bound), RenderConfig(), device=d)
fr = S.frame_to(S.make_frame(sub, mode="patch", n_patches=6, patch=32, seed=100), d)
vol = S.make_motion_weights_vol(sub.priors, 0).to(d)
rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).contiguous()
t_rand = torch.rand(rays.shape[0], 128, device=d)
z, x, mask = ops.warp_forward(rays, t_rand, fr.motion_scale_Rs.contiguous(), fr.motion_Ts.contiguous(), vol, fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz, 128)
xyz = x.reshape(-1, 3).contiguous(); m = xyz.shape[0]
st = net._static()
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(n)]; e1.record(); torch.cuda.synchronize(); return round(e0.elapsed_time(e1) / n, 4)
res = {"M": m, "mask_zero_frac": float((mask == 0).float().mean())}
idx = torch.empty(m, 4, 10, device=d, dtype=torch.int32)
def hier():
    ops.knn_hier(xyz, 128, *st["hier0"], idx, 0, 2, None, st["gid2"]); ops.knn_hier(xyz, 128, *st["hier1"], idx, 1, 3, st["gid1"], st["gid3"])
res["knn_hier_ms"] = t(hier)
ref = idx.clone()
for lr in (32, 16, 8, 4, 2, 1):
    res[f"knn_tree_lr{lr}_ms"] = t(lambda: ops.knn_tree(xyz, 128, st["tree"], lane_rays=lr))
    res[f"knn_tree_lr{lr}_equal"] = bool(torch.equal(ops.knn_tree(xyz, 128, st["tree"], lane_rays=lr), ref))
import time
torch.cuda.synchronize(); t0 = time.time()
grid = ops.build_knn_grid(st["point_base"], [f.to(d) for f in net.fps_index])
torch.cuda.synchronize(); res["grid_build_s"] = round(time.time() - t0, 2)
for cs in (0.02, 0.015):
    g2 = ops.build_knn_grid(st["point_base"], [f.to(d) for f in net.fps_index], cell=cs)
    res[f"knn_grid_cell{cs}_ms"] = t(lambda: ops.knn_grid(xyz, 128, g2, lane_rays=32))
    res[f"knn_grid_cell{cs}_MB"] = round((g2["entries"] * 2 + g2["cells"] * 32) / 2 ** 20, 1)
    res[f"knn_grid_cell{cs}_equal"] = bool(torch.equal(ops.knn_grid(xyz, 128, g2, lane_rays=32), ref))
    del g2
res["grid_cells"], res["grid_entries"], res["grid_MB"] = grid["cells"], grid["entries"], round((grid["entries"] * 2 + grid["cells"] * 24) / 2 ** 20, 1)
for lr in (32, 8):
    res[f"knn_grid_lr{lr}_ms"] = t(lambda: ops.knn_grid(xyz, 128, grid, lane_rays=lr))
    res[f"knn_grid_lr{lr}_equal"] = bool(torch.equal(ops.knn_grid(xyz, 128, grid, lane_rays=lr), ref))
ct = grid["cell_tab"][:, :, 1].float()
res["grid_mean_list_len"] = [round(float(ct[:, l].mean()), 1) for l in range(4)]
def chunked():
    for i in range(0, m, 300000):
        ops.knn_tree(xyz[i:i + 300000], 128, st["tree"], lane_rays=32)
res["knn_tree_chunked_ms"] = t(chunked)
flush = torch.empty(64 * 1024 * 1024, device=d)
def cold():
    flush.fill_(1.0); chunked()
res["knn_tree_chunked_cold_incl_flush_ms"] = t(cold)
res["flush_only_ms"] = t(lambda: flush.fill_(1.0))
print(json.dumps(res), flush=True)
# ---- hash grid backward per level
enc_in, _ = ops.sample_geometry(xyz, ref, st["point_base"], st["point_norms"], net.bound)
enc = net.cnl_mlp.module.encoder
import numpy as np
scales = ops.level_scales(float(np.log2(enc.per_level_scale)), enc.base_resolution, enc.num_levels, d)
gXB = torch.randn(m, 132, device=d) * 1e-3
g_emb = torch.zeros_like(enc.embeddings)
res2 = {}
for rl in (0, 8, 16):
    res2[f"hash_bwd_all_run{rl}_ms"] = t(lambda: ops.hashgrid_backward(gXB.data_ptr() + 4 * M.H_OFF, 132, 0, enc_in, enc.offsets, scales, g_emb, 2, run_length=rl))
xr = torch.rand(m, 4, device=d)
for rl in (0, 8):
    res2[f"hash_bwd_uniform_inputs_run{rl}_ms"] = t(lambda: ops.hashgrid_backward(gXB.data_ptr() + 4 * M.H_OFF, 132, 0, xr, enc.offsets, scales, g_emb, 2, run_length=rl))
res2["hash_fwd_all_ms"] = t(lambda: ops.hashgrid_forward(enc_in, enc.embeddings.detach(), enc.offsets, scales))
res2["hash_fwd_all_run16_ms"] = t(lambda: ops.hashgrid_forward(enc_in, enc.embeddings.detach(), enc.offsets, scales, run_length=16))
res2["hash_fwd_uniform_inputs_ms"] = t(lambda: ops.hashgrid_forward(xr, enc.embeddings.detach(), enc.offsets, scales))
g_mask = torch.randn(rays.shape[0], 128, device=d)
res2["warp_bwd_ms"] = t(lambda: ops.warp_backward(rays, t_rand, fr.motion_scale_Rs.contiguous(), fr.motion_Ts.contiguous(), fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz, g_mask, 128, tuple(vol.shape)))
res2["warp_fwd_ms"] = t(lambda: ops.warp_forward(rays, t_rand, fr.motion_scale_Rs.contiguous(), fr.motion_Ts.contiguous(), vol, fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz, 128))
print(json.dumps(res2), flush=True)
