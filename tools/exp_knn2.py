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
xyz = x.reshape(-1, 3).contiguous(); m = xyz.shape[0]
st = net._static()
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(n)]; e1.record(); torch.cuda.synchronize(); return round(e0.elapsed_time(e1) / n, 4)
res = {"persist": os.environ.get("OCCNERF_KNN_PERSIST", "1")}
for lo, hi in [(0, m), (0, 300000), (300000, 600000), (600000, m), (0, 303104), (0, 151552), (0, 131072), (131072, 262144), (0, 65536)]:
    q = xyz[lo:hi].contiguous()
    res[f"tree[{lo}:{hi}]"] = t(lambda: ops.knn_tree(q, 128, st["tree"], lane_rays=32))
# per-depth cost: all rays, 8 consecutive sample depths at a time
xs = x.reshape(-1, 128, 3)
for j0 in range(0, 128, 16):
    q = xs[:, j0:j0 + 16].reshape(-1, 3).contiguous()
    res[f"depth{j0}"] = t(lambda: ops.knn_tree(q, 16, st["tree"], lane_rays=32))
print(json.dumps(res))
