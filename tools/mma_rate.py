"""Cycles per tcgen05.mma (cta_group::1, M=128, K=16, bf16 SS) as a function of N and of how many SMs run at once."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import _lib
from occnerf_b200._lib import call, stream
d = torch.device("cuda")
_lib.load()
out = torch.zeros(256, device=d, dtype=torch.int64)
res = {}
for ctas in (1, 148):
    for n in (256, 128, 80, 16):
        for iters in (256, 2048):
            call("occnerf_mlp_debug_mma_rate", iters, n, out.data_ptr(), ctas, stream())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call("occnerf_mlp_debug_mma_rate", iters, n, out.data_ptr(), ctas, stream()); e1.record(); torch.cuda.synchronize()
            cyc = float(out[:ctas].double().mean()) / iters
            flop = 2.0 * 128 * n * 16
            res[f"ctas{ctas}_N{n}_iters{iters}"] = {"cycles_per_mma": round(cyc, 1), "flop_per_clk_per_sm": round(flop / cyc, 0),
                                                  "ms": round(e0.elapsed_time(e1), 4)}
print(json.dumps(res, indent=1))
