"""One launch each of the non-rigid chain (tf32, CTA pairs, 2^20 points) and of knn_grid (786 432 queries around the synthetic subject)
for an ncu capture: warm-ups first, the captured launches between cudaProfilerStart/Stop."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import ops, synthetic
d = torch.device("cuda")
torch.manual_seed(0)
sub = synthetic.make_subject()
w = synthetic.make_weights(sub.bound)
nr_w = [t.to(d) for t in w.nr_w]
nr_b = [t.to(d) for t in w.nr_b]
cond = torch.randn(1, 69, device=d) * 0.1
packed = ops.nonrigid_pack(nr_w, nr_b, cond, 2)
base = sub.point_base.to(d).float()
m = 1 << 20
xyz = (base[torch.randint(0, base.shape[0], (m,), device=d)] + torch.randn(m, 3, device=d) * 0.03).contiguous()
window = [1.0] * 6
grid = ops.build_knn_grid(base, [f.to(d) for f in sub.fps_index])
# queries ordered like samples along rays: 6144 rays x 128 depths through the body
N, S = 6144, 128
o = base[torch.randint(0, base.shape[0], (N,), device=d)] + torch.randn(N, 3, device=d) * 0.02
dirs = torch.nn.functional.normalize(torch.randn(N, 3, device=d), dim=-1)
t = torch.linspace(-1.0, 1.0, S, device=d)
q = (o[:, None, :] + dirs[:, None, :] * t[None, :, None]).reshape(-1, 3).contiguous()
for _ in range(3):
    ops.nonrigid_forward_tc(xyz, window, packed, 2)
    ops.knn_grid(q, S, grid)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.nonrigid_forward_tc(xyz, window, packed, 2)
ops.knn_grid(q, S, grid)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
