"""Micro-benchmark of the fused MLP forward kernel (forward-only and with activation saving)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import mlp as M, mlp_tc
from tests.test_mlp_gpu import _weights, _flat
d = torch.device("cuda")
m = int(os.environ.get("M", 262144))
once = "--once" in sys.argv
W = _flat(_weights(seed=2), d)
XB = torch.randn(m, 132, device=d) * 0.3
raw = torch.zeros(m, 5, device=d)
res = {}
for name, eng in [("tc1", mlp_tc.MlpTc(1)), ("tf32", mlp_tc.MlpTc(2)), ("tc3", mlp_tc.MlpTc(3)), ("tf32_cg1", mlp_tc.MlpTc(2, pair=False)), ("tc3_cg1", mlp_tc.MlpTc(3, pair=False))] + ([] if once else [("fp32", M.MlpSimt())]):
    for save in (False, True):
        if once and save: continue
        for _ in range(1 if once else 3):
            s = eng.forward(XB, raw, W, save=save); del s
        torch.cuda.synchronize()
        if once: continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            s = eng.forward(XB, raw, W, save=save); del s
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[f"{name}_save{int(save)}"] = dict(ms=ms, tflops=m * M.FLOP_FWD / ms / 1e9)
if not once:
    g_raw = torch.randn(m, 5, device=d)
    for name, eng in [("tc1", mlp_tc.MlpTc(1)), ("tf32", mlp_tc.MlpTc(2)), ("tc3", mlp_tc.MlpTc(3)), ("tf32_cg1", mlp_tc.MlpTc(2, pair=False)), ("tc3_cg1", mlp_tc.MlpTc(3, pair=False))]:
        saved = eng.forward(XB, raw, W, save=True)
        for _ in range(2): eng.backward(XB, g_raw, W, saved)
        torch.cuda.synchronize()
        from occnerf_b200 import _lib
        _lib.PROFILE = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): eng.backward(XB, g_raw, W, saved)
        e1.record(); torch.cuda.synchronize()
        prof, _lib.PROFILE = _lib.PROFILE, None
        res[f"{name}_bwd"] = dict(ms=e0.elapsed_time(e1) / 3, dgrad_ms=sum(a.elapsed_time(b) for a, b, _ in prof["occnerf_mlp_backward_tc"]) / 3)
print(json.dumps(res, indent=1))
