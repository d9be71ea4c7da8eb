"""Timing / numerics experiments on the tf32 MLP chain (occnerf_mlp_debug_set): weight-stream hypotheses.

  half    : the pair producer fetches half of every weight chunk (garbage results; is the W wait L2-bound?)
  copiesN : N replicas of the packed weight image, pair p streams replica p % N (same-address hot spot in L2?)
  nomask  : tf32 rounding as one integer add, low 13 mantissa bits left in place (does the tensor core ignore them?)
"""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import mlp as M, mlp_tc, _lib
from tests.test_mlp_gpu import _weights, _flat
d = torch.device("cuda")
m = int(os.environ.get("M", 262144))
W = _flat(_weights(seed=2), d)
XB = torch.randn(m, 132, device=d) * 0.3
g_raw = torch.randn(m, 5, device=d)
lib = _lib.load()
buf = (ctypes.c_ulonglong * 16)()


def counters():
    lib.occnerf_mlp_debug_counters(ctypes.cast(buf, ctypes.c_void_p), 1)
    v = list(buf)
    return {"mma_wait_W_pct": round(100 * v[0] / max(v[4], 1), 1), "mma_wait_A_pct": round(100 * v[1] / max(v[4], 1), 1),
            "epi_wait_acc_pct": round(100 * v[2] / max(v[5], 1), 1)}


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
ref_raw = None
for name, exp, copies in (("base", 0, 1), ("half", 1, 1), ("copies2", 0, 2), ("copies4", 0, 4), ("copies8", 0, 8), ("copies16", 0, 16),
                          ("nomask", 4, 1)):
    eng = mlp_tc.MlpTc(2)
    eng.copies = copies
    lib.occnerf_mlp_debug_set(0, exp, copies)
    raw = torch.zeros(m, 5, device=d)
    r = {}
    for save in (False, True):
        r[f"fwd_save{int(save)}_ms"] = round(timed(lambda: eng.forward(XB, raw, W, save=save)), 4)
    saved = eng.forward(XB, raw, W, save=True)
    torch.cuda.synchronize()
    _lib.PROFILE = {}
    for _ in range(4):
        eng.backward(XB, g_raw, W, saved)
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    r["dgrad_ms"] = round(sum(a.elapsed_time(b) for a, b, _ in prof["occnerf_mlp_backward_tc"][1:]) / 3, 4)
    if name == "base":
        ref_raw = raw.clone()
    elif exp != 1:
        r["max_abs_diff_vs_base"] = float((raw - ref_raw).abs().max())
        r["bitwise_equal"] = bool(torch.equal(raw, ref_raw))
    # stall counters of one forward (save=1)
    lib.occnerf_mlp_debug_set(1, exp, copies)
    eng.forward(XB, raw, W, save=True); counters()
    eng.forward(XB, raw, W, save=True); r["stalls_fwd_save1"] = counters()
    # per-layer timeline of one tile (save=1): layer period, MMA issue span, W / A waits of the MMA thread, epilogue span
    lib.occnerf_mlp_debug_set(17, exp, copies)
    for _ in range(2):
        eng.forward(XB, raw, W, save=True)
    tb = (ctypes.c_ulonglong * 192)()
    lib.occnerf_mlp_debug_trace(ctypes.cast(tb, ctypes.c_void_p))
    v = [list(tb[l * 12:(l + 1) * 12]) for l in range(10)]
    r["trace_layers_1_3"] = [dict(period=v[l + 1][0] - v[l][0], issue_span=v[l][2] - v[l][1], wait_W=v[l][11], wait_A=v[l][10],
                                  acc_ready_after_last_issue=v[l][3] - v[l][2], first_pub_after_acc=v[l][4] - v[l][3],
                                  epi_span=v[l][5] - v[l][3]) for l in (1, 2, 3)]
    lib.occnerf_mlp_debug_set(0, 0, 1)
    res[name] = r
    del saved
print(json.dumps(res, indent=1))
