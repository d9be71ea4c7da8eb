"""Weight-stream round trip of the pair-mode tf32 forward chain: per chunk of layers 2 and 3 of one tile, when the producer saw the
slot empty / issued the copy, and when the MMA thread started waiting / saw its half / saw the peer's half / had issued the MMAs
(clock64 of the leader CTA's SM, relative to the first event).  OCCNERF_MLP_DEBUG=17 instrumentation, csrc/mlp_tc.cu g_trace_w."""
import sys, os, ctypes
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
lib.occnerf_mlp_debug_set(17)
names = ["prod_slot_empty", "prod_issued", "mma_wait_start", "own_half", "peer_half", "mma_issued"]
for npass, name in ((2, "tf32"), (1, "tc1")):
    eng = mlp_tc.MlpTc(npass)
    for save in (False, True):
        for _ in range(2):
            eng.forward(XB, raw, W, save=save)
        buf = (ctypes.c_ulonglong * 96)()
        lib.occnerf_mlp_debug_trace_w(ctypes.cast(buf, ctypes.c_void_p))
        v = list(buf)
        t0 = min(x for x in v if x)
        print(f"== {name} save={int(save)}")
        print("layer chunk " + " ".join(f"{n:>16}" for n in names))
        for l in range(2):
            for c in range(8):
                row = v[(l * 8 + c) * 6:(l * 8 + c + 1) * 6]
                print(f"{l + 2:5d} {c:5d} " + " ".join(f"{(x - t0) if x else 0:16d}" for x in row))
