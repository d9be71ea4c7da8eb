#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python tools/sweep_r2.py --max-log2 20 > gpurun_out/r2k_sweep_1gpu.md 2> gpurun_out/r2k_sweep_1gpu.err
tail -n 3 gpurun_out/r2k_sweep_1gpu.err
timeout 600 python tools/freeview_bench.py --views 10 --res 1024 > gpurun_out/r2k_freeview_1gpu.json 2> gpurun_out/r2k_freeview_1gpu.err
cat gpurun_out/r2k_freeview_1gpu.json
timeout 600 python bench.py --workload ocmotion --steps 10 --warmup 3 > gpurun_out/r2k_bench_ocmotion.json 2> gpurun_out/r2k_bench_ocmotion.err
timeout 900 python bench.py --impl reference --ref-rays 1536 --steps 1 --warmup 1 > gpurun_out/r2k_ref_1536.json 2> gpurun_out/r2k_ref_1536.err
timeout 300 python bench.py --impl reference --ref-rays 96 --steps 2 --warmup 1 > gpurun_out/r2k_ref_96.json 2> gpurun_out/r2k_ref_96.err
cat gpurun_out/r2k_ref_1536.json gpurun_out/r2k_ref_96.json | cut -c 1-400
timeout 600 ncu --set full --clock-control none -k regex:"rays_|image_" -c 10 -o gpurun_out/r2k_rays_image python tools/freeview_bench.py --views 1 --warmup 1 --res 1024 > gpurun_out/r2k_ncu_rays.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench_ocmotion.json'))
print('ocmotion', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
PY
