#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_train_step_gpu.py -x -q -m gpu > gpurun_out/r2h_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2h_tests.log
tail -n 30 gpurun_out/r2h_tests.log
timeout 900 python bench.py --engine tf32 --steps 10 --warmup 3 > gpurun_out/r2h_bench_tf32.json 2> gpurun_out/r2h_bench_tf32.err
tail -n 5 gpurun_out/r2h_bench_tf32.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_tf32.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d.get('cpu_baseline'))
print([(k['call'][8:], round(k['ms_per_step'],3)) for k in d['kernels'][:14]])
print(d['forward_only'])
PY
