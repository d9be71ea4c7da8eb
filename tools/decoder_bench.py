"""Motion-weight volume decoder, forward + backward per step: native kernels (csrc/deconv.cu) vs the library path (cuDNN), both tf32."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import prologue as P, _lib
d = torch.device("cuda")
torch.manual_seed(0)
dec = P.MotionWeightVolumeDecoder().to(d)
priors = torch.rand(1, 25, 32, 32, 32, device=d) + 0.01
gv = torch.randn(1, 25, 32, 32, 32, device=d)
res = {}
def step():
    dec.zero_grad(set_to_none=True)
    (dec(motion_weights_priors=priors) * gv).sum().backward()
for name, native, overlap in (("library_cudnn_tf32", False, 1), ("native_tf32_one_stream", True, 0), ("native_tf32", True, 1)):
    dec.native = native
    _lib.call("occnerf_deconv_set_overlap", overlap)   # weight gradients on the side stream beside the data gradients
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        step()
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    res[name + "_fwd_bwd_ms"] = e0.elapsed_time(e1) / 10
if True:
    dec.native = True
    _lib.PROFILE = {}
    for _ in range(5): step()
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    for k, v in prof.items():
        res["call:" + k] = [round(a.elapsed_time(b), 4) for a, b, _ in v[-(len(v) // 5):]]
print(json.dumps(res, indent=1))
