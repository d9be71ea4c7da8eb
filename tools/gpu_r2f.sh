#!/usr/bin/env bash
mkdir -p gpurun_out
OCCNERF_MLP_PAIR=1 timeout 300 python -m pytest tests/test_mlp_gpu.py -x -q -m gpu > gpurun_out/r2f_tests_pair.log 2>&1
echo "tests exit $?" >> gpurun_out/r2f_tests_pair.log
tail -n 25 gpurun_out/r2f_tests_pair.log
timeout 300 python tools/bench_mlp.py > gpurun_out/r2f_bench_mlp.json 2> gpurun_out/r2f_bench_mlp.err
OCCNERF_MLP_PAIR=1 timeout 200 python tools/mlp_trace.py > gpurun_out/r2f_mlp_trace.txt 2> gpurun_out/r2f_mlp_trace.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_mlp.json'))
print({k:round(v['ms'],4) for k,v in d.items()})
PY
