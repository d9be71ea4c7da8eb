import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import ops, synthetic as S
from occnerf_b200.network import RenderConfig
d = torch.device("cuda")
sub = S.make_subject(0)
net = S.network_from_synthetic(sub, S.make_weights(sub.bound), RenderConfig(), device=d)
fr = S.frame_to(S.make_frame(sub, mode="patch", n_patches=6, patch=32, seed=100), d)
vol = S.make_motion_weights_vol(sub.priors, 0).to(d)
rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).contiguous()
t_rand = torch.rand(rays.shape[0], 128, device=d)
z, x, mask = ops.warp_forward(rays, t_rand, fr.motion_scale_Rs.contiguous(), fr.motion_Ts.contiguous(), vol, fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz, 128)
xyz = x.reshape(-1, 3).contiguous()
st = net._static()
grid = ops.build_knn_grid(st["point_base"], [f.to(d) for f in net.fps_index])
idx = ops.knn_grid(xyz, 128, grid).view(6144, 128, 4, 10)
res = {}
a, b = idx[:, :-1], idx[:, 1:]
for lev in range(4):
    res[f"same_slot_l{lev}"] = round(float((a[:, :, lev] == b[:, :, lev]).float().mean()), 3)
    member = (a[:, :, lev, :, None] == b[:, :, lev, None, :]).any(-1).float().mean()
    res[f"member_of_next_l{lev}"] = round(float(member), 3)
    allsame = (a[:, :, lev] == b[:, :, lev]).all(-1).float().mean()
    res[f"whole_list_same_l{lev}"] = round(float(allsame), 3)
print(json.dumps(res))
