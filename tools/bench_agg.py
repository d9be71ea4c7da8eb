import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import ops, synthetic as S
d = torch.device("cuda")
sub = S.make_subject(0)
from occnerf_b200.network import RenderConfig
net = S.network_from_synthetic(sub, S.make_weights(sub.bound), RenderConfig(), device=d)
fr = S.frame_to(S.make_frame(sub, mode="patch", n_patches=2, patch=32, seed=100), d)
vol = S.make_motion_weights_vol(sub.priors, 0).to(d)
rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).contiguous()
z, x, m = ops.warp_forward(rays, None, fr.motion_scale_Rs.contiguous(), fr.motion_Ts.contiguous(), vol, fr.cnl_bbox_min_xyz, fr.cnl_bbox_scale_xyz, 128)
xyz = x.reshape(-1, 3).contiguous(); M = xyz.shape[0]
st = net._static()
idx = torch.empty(M, 4, 10, device=d, dtype=torch.int32)
ops.knn_hier(xyz, 128, *st["hier0"], idx, 0, 2, None, st["gid2"]); ops.knn_hier(xyz, 128, *st["hier1"], idx, 1, 3, st["gid1"], st["gid3"])
gX = torch.randn(M, 132, device=d); counter = torch.ones(6890, device=d)
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(n)]; e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
res = {"M": M}
for c in (1, 4, 16, 64):
    res[f"v1_copies{c}"] = t(lambda: ops.aggregate_backward(idx, counter, gX.data_ptr() + 256, 132, 6890, copies=c))
res["v2"] = t(lambda: ops.aggregate_backward(idx, counter, gX.data_ptr() + 256, 132, 6890, group_stride=128))
print(json.dumps(res))
