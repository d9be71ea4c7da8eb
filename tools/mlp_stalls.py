"""Where the role threads of the fused MLP chain kernels wait (OCCNERF_MLP_DEBUG=1 instrumentation)."""
import sys, os, json, ctypes
os.environ["OCCNERF_MLP_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import mlp as M, mlp_tc, _lib
from tests.test_mlp_gpu import _weights, _flat
d = torch.device("cuda")
m = int(os.environ.get("M", 262144))
W = _flat(_weights(seed=2), d)
XB = torch.randn(m, 132, device=d) * 0.3
raw = torch.zeros(m, 5, device=d)
g_raw = torch.randn(m, 5, device=d)
lib = _lib.load()
buf = (ctypes.c_ulonglong * 16)()
def read(reset=1):
    lib.occnerf_mlp_debug_counters(ctypes.cast(buf, ctypes.c_void_p), reset)
    v = list(buf)
    ctas = max(v[6], 1)
    return {"mma_wait_weights_pct": round(100 * v[0] / max(v[4], 1), 1), "mma_wait_A_pct": round(100 * v[1] / max(v[4], 1), 1),
            "epi_wait_acc_pct": round(100 * v[2] / max(v[5], 1), 1), "epi_tmem_ld_wait_pct": round(100 * v[8] / max(v[5], 1), 1),
            "epi_publish_pct": round(100 * v[9] / max(v[5], 1), 1), "producer_wait_slot_cycles_per_cta": v[3] // ctas,
            "mma_thread_cycles_per_cta": v[4] // ctas, "epi_thread_cycles_per_cta": v[5] // ctas, "ctas": v[6]}
res = {}
for name, eng in [("tc1", mlp_tc.MlpTc(1)), ("tf32", mlp_tc.MlpTc(2)), ("tc3", mlp_tc.MlpTc(3))]:
    for save in (False, True):
        s = eng.forward(XB, raw, W, save=save)
        read()
        s = eng.forward(XB, raw, W, save=save)
        res[f"{name}_fwd_save{int(save)}"] = read()
    read()
    eng.backward(XB, g_raw, W, s)
    res[f"{name}_bwd(dgrad)"] = read()
print(json.dumps(res, indent=1))
