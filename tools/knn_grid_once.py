import sys, os
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
xyz = x.reshape(-1, 3)[:262144].contiguous()
st = net._static()
grid = ops.build_knn_grid(st["point_base"], [f.to(d) for f in net.fps_index])
ops.knn_grid(xyz, 128, grid)
ops.knn_grid(xyz, 128, grid)
torch.cuda.synchronize()
