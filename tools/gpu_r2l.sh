#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prologue_gpu.py tests/test_train_step_gpu.py tests/test_view_gpu.py -x -q -m gpu > gpurun_out/r2l_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2l_tests.log
tail -n 15 gpurun_out/r2l_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['our_launches_per_step'])
PY
