import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from occnerf_b200 import ops, synthetic as S
from oracle import make_golden, occnerf_oracle as O
from tests.helpers import load_case
from tests.test_render_gpu import _net, _render

name = sys.argv[1] if len(sys.argv) > 1 else "train_init_early"
sub, w, fr, vol, t_rand, rk, g = load_case(name)
from tests.test_oracle_golden import _params_with_grad
sub0, w0 = sub, w
sub, w = _params_with_grad(sub, w)
d = torch.device("cuda")
if not torch.cuda.is_available(): print("cpu only")
# oracle grads wrt raw/mask
volr = vol.clone().requires_grad_(True)
o = O.render_rays(fr, volr, sub, w, iter_val=rk["iter_val"], training=True, t_rand=t_rand, return_aux=True)
make_golden.scalar_loss(o).backward()
raw_o, mask_o = o["raw"].detach().requires_grad_(True), o["mask"].detach().requires_grad_(True)
rgb2, acc2, depth2, term2, _ = O.composite(raw_o, mask_o, o["z"], fr.rays_d, fr.bgcolor)
make_golden.scalar_loss(dict(rgb=rgb2, alpha=acc2, depth=depth2, comp_loss=O.completeness_term(raw_o))).backward()
print("oracle g_raw", raw_o.grad.abs().max(0)[0].max(0)[0], "g_mask", mask_o.grad.abs().max())
N = fr.rays_o.shape[0]
rays = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).to(d)
rgbw = torch.tensor(make_golden.LOSS_W["rgb"]).expand(N, 3).contiguous().to(d)
g_acc = torch.full((N,), 0.5, device=d); g_depth = torch.full((N,), 0.25, device=d)
g_comp = torch.full((N, 128), 1.0 / (N * 128), device=d)
g_raw, g_mask = ops.composite_backward(raw_o.detach().to(d).contiguous(), mask_o.detach().to(d).contiguous(), o["z"].to(d).contiguous(), rays.contiguous(),
                                       fr.bgcolor.to(d), rgbw, g_acc, g_depth, g_comp)
for c in range(5):
    print("g_raw ch", c, float((g_raw[..., c].cpu() - raw_o.grad[..., c]).abs().max()), float(raw_o.grad[..., c].abs().max()))
print("g_mask", float((g_mask.cpu() - mask_o.grad).abs().max()), float(mask_o.grad.abs().max()))
# e2e
net = _net(sub0, w0, rk)
vol_d = vol.to(d).requires_grad_(True)
out = _render(net, fr, vol_d, t_rand, rk["iter_val"])
make_golden.scalar_loss({k: out[k] for k in ("rgb", "alpha", "depth", "comp_loss")}).backward()
print("e2e g_vol", float((vol_d.grad.cpu() - volr.grad).abs().max()), float(volr.grad.abs().max()))
m = net.cnl_mlp.module
def rel(a, b): return float((a.cpu() - b).abs().max() / b.abs().max())
print("geo_w", rel(m.geo_linear[0].weight.grad, w.geo_w.grad), "geo_b", rel(m.geo_linear[0].bias.grad, w.geo_b.grad))
print("geo_w row0", rel(m.geo_linear[0].weight.grad[0], w.geo_w.grad[0]), "rows1+", rel(m.geo_linear[0].weight.grad[1:], w.geo_w.grad[1:]))
for i, li in enumerate((0, 2, 4, 6)):
    print("pts", i, rel(m.pts_linears[li].weight.grad, w.pts_w[i].grad), "rgb", i, rel(m.rgb_linears[li].weight.grad, w.rgb_w[i].grad))
print("emb", rel(m.encoder.embeddings.grad, w.embeddings.grad), "point_dist", rel(net.point_dist.grad, sub.point_dist.grad))
