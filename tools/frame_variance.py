"""Single-GPU step time of the eight per-rank workloads of `bench.py --gpus 8` (each rank renders its own frame: seed 100 + rank), one
after the other on one GPU: separates workload imbalance between ranks from the cost of the gradient all-reduce in the 8-GPU number."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from occnerf_b200 import _lib
_lib.load()
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
res = {}
for r in range(8):
    wl = bench.Workload(dev, r, "tf32")
    for _ in range(3):
        wl.step_device(1)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        wl.step_device(1)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        wl.step_device(1)
    res[f"rank{r}_ms"] = round(bench.timed_loop(lambda: g.replay(), 10, 3, 1, flush), 4)
    del g, wl
    torch.cuda.empty_cache()
res["max_ms"] = max(v for v in res.values())
res["rank0_ms"] = res["rank0_ms"]
print(json.dumps(res, indent=1))
