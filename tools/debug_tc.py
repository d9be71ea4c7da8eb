import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import mlp as M, mlp_tc
from tests.test_mlp_gpu import _weights, _flat, _inputs
d = torch.device("cuda")
m = 3000
w = _weights(seed=2); agg, var, h = _inputs(m, seed=5)
W = _flat(w, d)
def run(eng):
    XB = torch.zeros(m, 132, device=d)
    XB[:, 64:99], XB[:, 99:100], XB[:, 100:] = agg.to(d), var.to(d), h.to(d)
    raw = torch.zeros(m, 5, device=d)
    saved = eng.forward(XB, raw, W, save=True)
    g_raw = torch.zeros(m, 5, device=d); g_raw[:, :4] = torch.randn(m, 4, generator=torch.Generator().manual_seed(1)).to(d)
    gXB, grads = eng.backward(XB, g_raw, W, saved)
    return XB, raw, saved, gXB, grads
a = run(M.MlpSimt()); b = run(mlp_tc.MlpTc(3))
print("XB", (a[0]-b[0]).abs().max().item(), "raw", (a[1]-b[1]).abs().max().item())
for i in range(8):
    x, y = a[2]["acts"][i], b[2]["acts"][i]
    print("act", i, (x-y).abs().max().item(), x.abs().max().item(), "mask flips", ((x>0)!=(y>0)).sum().item(), "rows bad", ((x-y).abs().max(1)[0] > 1e-3).sum().item())
print("gXB", (a[3]-b[3]).abs().max().item(), a[3].abs().max().item())
for n, x, y in zip(M.MlpWeights.ORDER, a[4], b[4]):
    print(n, ((x-y).abs().max()/x.abs().max()).item())
