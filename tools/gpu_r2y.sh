#!/usr/bin/env bash
# ncu --set full of the native decoder kernels (one forward + backward), summarised on the box
mkdir -p gpurun_out
cat > /tmp/dec_once.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from occnerf_b200 import prologue as P
d = torch.device("cuda"); torch.manual_seed(0)
dec = P.MotionWeightVolumeDecoder().to(d); dec.native = True
priors = torch.rand(1, 25, 32, 32, 32, device=d) + 0.01
gv = torch.randn(1, 25, 32, 32, 32, device=d)
for _ in range(2):
    dec.zero_grad(set_to_none=True)
    (dec(motion_weights_priors=priors) * gv).sum().backward()
torch.cuda.synchronize()
torch.cuda.profiler.start()
dec.zero_grad(set_to_none=True)
(dec(motion_weights_priors=priors) * gv).sum().backward()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
PY
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'deconv_|linear_|bias_|weight_volume' -c 40 -f -o /tmp/r2y_deconv python /tmp/dec_once.py > gpurun_out/r2y_deconv_ncu.out 2>&1
python tools/summarize_ncu.py full /tmp/r2y_deconv.ncu-rep gpurun_out/r2y_deconv_kernels_full.md
ls -la gpurun_out/r2y_*
