#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python tools/mma_rate.py > gpurun_out/r2c_mma_rate.json 2> gpurun_out/r2c_mma_rate.err
timeout 200 python tools/mlp_trace.py > gpurun_out/r2c_mlp_trace.txt 2> gpurun_out/r2c_mlp_trace.err
timeout 300 python tools/diag_point_dist.py > gpurun_out/r2c_diag_pd.log 2>&1
cat gpurun_out/r2c_mma_rate.json | head -60
