"""Per-layer timeline of one tile of the fused MLP forward chain (OCCNERF_MLP_DEBUG=17: counters + trace)."""
import sys, os, json, ctypes
os.environ.setdefault("OCCNERF_MLP_DEBUG", "17")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import mlp as M, mlp_tc, _lib
from tests.test_mlp_gpu import _weights, _flat
d = torch.device("cuda")
m = 262144
W = _flat(_weights(seed=2), d)
XB = torch.randn(m, 132, device=d) * 0.3
raw = torch.zeros(m, 5, device=d)
lib = _lib.load()
buf = (ctypes.c_ulonglong * 192)()
names = ["mma_enter", "mma_first_issue", "mma_last_issue", "epi0_acc_ready", "epi0_first_pub", "epi0_last_pub", "epi15_acc_ready", "epi15_last_pub",
         "prod_last_chunk", "-", "mma_wait_A_cyc", "mma_wait_W_cyc"]
g_raw = torch.randn(m, 5, device=d)
for name, npass in (("tc1", 1), ("tf32", 2), ("tc3", 3)):
    eng = mlp_tc.MlpTc(npass)
    for save in (False, True, "dgrad"):
        if save == "dgrad":
            saved = eng.forward(XB, raw, W, save=True)
            for _ in range(2):
                eng.backward(XB, g_raw, W, saved)
            del saved
        else:
            for _ in range(2):
                eng.forward(XB, raw, W, save=save)
        lib.occnerf_mlp_debug_trace(ctypes.cast(buf, ctypes.c_void_p))
        v = [list(buf[l * 12:(l + 1) * 12]) for l in range(10)]
        t0 = v[0][0]
        print(f"== {name} save={save if save == 'dgrad' else int(save)}  (cycles relative to the MMA thread entering layer 0 of the tile)")
        print("layer " + " ".join(f"{n:>16}" for n in names))
        for l in range(10):
            row = [(x - t0) if i < 9 and x else x for i, x in enumerate(v[l])]
            print(f"{l:5d} " + " ".join(f"{x:16d}" for x in row))
