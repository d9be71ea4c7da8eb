#!/usr/bin/env bash
# round-2 closing evidence: the default bench line (with the CPU baseline), the reference arm, the ncu launch list of two steady-state steps
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2z_bench_1gpu.out 2> gpurun_out/r2z_bench_1gpu.err
tail -1 gpurun_out/r2z_bench_1gpu.out | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2z_launches.csv \
    python bench.py --profile-mode --steps 2 --warmup 3 > gpurun_out/r2z_launches.out 2>&1
python tools/summarize_ncu.py launches gpurun_out/r2z_launches.csv gpurun_out/r2z_launches.md
head -12 gpurun_out/r2z_launches.md
