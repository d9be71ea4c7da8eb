#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_gpu.py tests/test_render_gpu.py -x -q -m gpu > gpurun_out/r2d_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2d_tests.log
timeout 300 python tools/bench_mlp.py > gpurun_out/r2d_bench_mlp.json 2> gpurun_out/r2d_bench_mlp.err
timeout 200 python tools/mlp_trace.py > gpurun_out/r2d_mlp_trace.txt 2> gpurun_out/r2d_mlp_trace.err
tail -n 4 gpurun_out/r2d_tests.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_mlp.json'))
print({k:round(v['ms'],4) for k,v in d.items()})
PY
