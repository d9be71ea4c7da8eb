#!/usr/bin/env bash
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2g_tests.log
tail -n 5 gpurun_out/r2g_tests.log
timeout 600 python bench.py --engine tf32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_tf32.json 2> gpurun_out/r2g_bench_tf32.err
timeout 600 python bench.py --engine tc3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_tc3.json 2> gpurun_out/r2g_bench_tc3.err
python - <<'PY'
import json
for e in ("tf32","tc3"):
    d=json.load(open(f'gpurun_out/r2g_bench_{e}.json'))
    print(e, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])
    print([(k['call'][8:], round(k['ms_per_step'],3)) for k in d['kernels'][:12]])
PY
