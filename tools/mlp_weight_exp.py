"""Timing experiment on the weight stream of the fused MLP chain (OCCNERF_MLP_DEBUG bits, csrc/mlp_tc.cu producer_loop):
run once per setting, e.g.  OCCNERF_MLP_DEBUG=1 (baseline with counters) / 3 (half the bytes) / 5 (own half, no multicast) /
9 (whole chunk per CTA, no multicast)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import mlp as M, mlp_tc
from tests.test_mlp_gpu import _weights, _flat
d = torch.device("cuda")
m = 262144
W = _flat(_weights(seed=2), d)
XB = torch.randn(m, 132, device=d) * 0.3
raw = torch.zeros(m, 5, device=d)
res = {"debug": os.environ.get("OCCNERF_MLP_DEBUG", "0")}
for name, eng in [("tc1", mlp_tc.MlpTc(1)), ("tf32", mlp_tc.MlpTc(2)), ("tc3", mlp_tc.MlpTc(3))]:
    for _ in range(3):
        eng.forward(XB, raw, W, save=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.forward(XB, raw, W, save=False)
    e1.record(); torch.cuda.synchronize()
    res[name + "_fwd_ms"] = round(e0.elapsed_time(e1) / 5, 4)
print(json.dumps(res))
