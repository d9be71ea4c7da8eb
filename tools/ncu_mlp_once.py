"""One tf32 forward chain launch (save=True) for an ncu capture (warm-ups first; the profiled launch is selected with --launch-skip)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import mlp_tc
from tests.test_mlp_gpu import _weights, _flat
d = torch.device("cuda")
m = int(os.environ.get("M", 262144))
W = _flat(_weights(seed=2), d)
XB = torch.randn(m, 132, device=d) * 0.3
raw = torch.zeros(m, 5, device=d)
eng = mlp_tc.MlpTc(int(os.environ.get("NPASS", 2)))
for _ in range(6):
    s = eng.forward(XB, raw, W, save=True)
torch.cuda.synchronize()
