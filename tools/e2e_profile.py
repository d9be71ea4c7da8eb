"""Kernel-level breakdown of one e2e step (Network.forward with the library prologue + loss + backward + Adam)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from occnerf_b200 import _lib
_lib.load()
dev = torch.device("cuda", 0)
wl = bench.Workload(dev, 0, os.environ.get("ENGINE", "tf32"))
MODE = os.environ.get("MODE", "e2e")
step = (lambda: bench.Workload.step_e2e(wl, 1)) if MODE == "e2e" else (lambda: wl.step_device(1))
wl.step_e2e = lambda _w: step()
for _ in range(3):
    wl.step_e2e(1)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        wl.step_e2e(1)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t > 0 and e.device_type.name == "CUDA":
        rows.append((t / 3 / 1000.0, e.count / 3, e.key[:90]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total device ms/step", round(tot, 3))
for ms, n, k in rows[:40]:
    print(f"{ms:8.3f} ms  x{n:5.1f}  {k}")
import time
t0 = time.time()
for _ in range(5):
    wl.step_e2e(1)
torch.cuda.synchronize()
print("wall ms/step", (time.time() - t0) / 5 * 1e3)
