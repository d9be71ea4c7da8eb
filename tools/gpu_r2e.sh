#!/usr/bin/env bash
mkdir -p gpurun_out
M=262144 timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp_chain_tc_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/r2e_mlp_tf32_fwd python tools/ncu_mlp_once.py > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out/r2e_mlp_tf32_fwd.ncu-rep
