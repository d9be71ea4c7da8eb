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
z, x, m = ops.warp_forward(rays, None, fr.motion_scale_Rs.contiguous(), fr.motion_Ts.contiguous(), vol, fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz, 128)
xyz = x.reshape(-1, 3)[:300000].contiguous(); M = xyz.shape[0]
st = net._static()
grid = ops.build_knn_grid(st["point_base"], [f.to(d) for f in net.fps_index])
idx = ops.knn_grid(xyz, 128, grid)
gX = torch.randn(M, 132, device=d); counter = torch.ones(6890, device=d)
X = torch.empty(M, 132, device=d); f36 = torch.randn(6890, 36, device=d)
gp = torch.zeros(64, 6890, 36, device=d)
for rep in range(2):
    ops.aggregate_forward(idx, counter, f36, X.data_ptr() + 256, 132)
    aw = ops.aggregate_forward(idx, counter, f36, X.data_ptr() + 256, 132, want_att=True)
    ops.aggregate_backward(idx, counter, gX.data_ptr() + 256, 132, 6890, g_priv=gp)
    ops.aggregate_backward(idx, counter, gX.data_ptr() + 256, 132, 6890, g_priv=gp, att_w=aw)
torch.cuda.synchronize()
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(n)]; e1.record(); torch.cuda.synchronize(); return round(e0.elapsed_time(e1) / n, 4)
aw0 = torch.empty(M, 40, device=d)
print({"fwd": t(lambda: ops.aggregate_forward(idx, counter, f36, X.data_ptr() + 256, 132)),
       "fwd_att": t(lambda: ops.aggregate_forward(idx, counter, f36, X.data_ptr() + 256, 132, want_att=True)),
       "bwd": t(lambda: ops.aggregate_backward(idx, counter, gX.data_ptr() + 256, 132, 6890, g_priv=gp)),
       "bwd_slot": t(lambda: ops.aggregate_backward(idx, counter, gX.data_ptr() + 256, 132, 6890, g_priv=gp, att_w=aw))})
