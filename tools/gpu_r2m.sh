#!/usr/bin/env bash
# 8-GPU evidence: data-parallel training bench, freeview (configs[2]) at 2/4/8, whole-path sweep at 8
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 600 bash -c "$(declare -f run); run 8 29521 bench.py --gpus 8 --steps 20 --warmup 5" > gpurun_out/r2m_bench_8gpu.json 2> gpurun_out/r2m_bench_8gpu.err
timeout 600 bash -c "$(declare -f run); run 4 29522 bench.py --gpus 4 --steps 20 --warmup 5" > gpurun_out/r2m_bench_4gpu.json 2> gpurun_out/r2m_bench_4gpu.err
timeout 600 bash -c "$(declare -f run); run 8 29523 bench.py --gpus 8 --steps 10 --warmup 3 --workload ocmotion" > gpurun_out/r2m_bench_ocmotion_8gpu.json 2> gpurun_out/r2m_bench_ocmotion_8gpu.err
for n in 2 4 8; do
  timeout 600 bash -c "$(declare -f run); run $n 2953$n tools/freeview_bench.py --views 10 --res 1024" > gpurun_out/r2m_freeview_${n}gpu.json 2> gpurun_out/r2m_freeview_${n}gpu.err
done
timeout 900 bash -c "$(declare -f run); run 8 29541 tools/sweep_r2.py --min-log2 18 --max-log2 22 --samples 128" > gpurun_out/r2m_sweep_8gpu.md 2> gpurun_out/r2m_sweep_8gpu.err
python - <<'PY'
import json
for f in ("bench_8gpu","bench_4gpu","bench_ocmotion_8gpu","freeview_2gpu","freeview_4gpu","freeview_8gpu"):
    try:
        for line in open(f'gpurun_out/r2m_{f}.json'):
            if line.startswith('{'):
                d=json.loads(line); print(f, d['value'], d.get('ms_per_step', d.get('ms_per_view')), d.get('e2e',{}).get('value'), d.get('e2e',{}).get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
head -n 12 gpurun_out/r2m_sweep_8gpu.md | cut -c 1-200
