#!/usr/bin/env bash
# round-2 ncu evidence: launch list of two steady-state steps + one --set full capture of every kernel of ours in one step.
# The .ncu-rep is summarised on the box (tools/summarize_ncu.py) and removed: gpurun brings back at most 64 MiB.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2x_launches.csv \
    python bench.py --profile-mode --steps 2 --warmup 3 > gpurun_out/r2x_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off \
    -k regex:'mlp_|hashgrid_|aggregate_|knn_|warp_|composite_|vertex_|clip_adam|grad_sumsq|sample_geometry|pack_|unpack_' -c 50 -f -o /tmp/r2x_kernels \
    python bench.py --profile-mode --steps 1 --warmup 3 > gpurun_out/r2x_kernels.out 2>&1
python tools/summarize_ncu.py full /tmp/r2x_kernels.ncu-rep gpurun_out/r2x_kernels_full.md
python tools/summarize_ncu.py launches gpurun_out/r2x_launches.csv gpurun_out/r2x_launches.md
ls -la /tmp/r2x_kernels.ncu-rep gpurun_out/r2x_*
