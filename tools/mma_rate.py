"""Cycles per tcgen05.mma (cta_group::1, M=128, SS operands; kind::f16 bf16 K=16 and kind::tf32 K=8) as a function of N and
of how many SMs run at once."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import _lib
from occnerf_b200._lib import call, stream
d = torch.device("cuda")
_lib.load()
out = torch.zeros(256, device=d, dtype=torch.int64)
res = {}
for tf32 in (0, 1):
    for ctas in (1, 148):
        for n in (256, 128, 80, 16):
            iters = 2048
            call("occnerf_mlp_debug_mma_rate", iters, n, tf32, out.data_ptr(), ctas, stream())
            torch.cuda.synchronize()
            call("occnerf_mlp_debug_mma_rate", iters, n, tf32, out.data_ptr(), ctas, stream())
            torch.cuda.synchronize()
            cyc = float(out[:ctas].double().mean()) / iters
            flop = 2.0 * 128 * n * (8 if tf32 else 16)
            res[f"{'tf32' if tf32 else 'bf16'}_ctas{ctas}_N{n}"] = {"cycles_per_mma": round(cyc, 1), "flop_per_clk_per_sm": round(flop / cyc, 0)}
print(json.dumps(res, indent=1))
