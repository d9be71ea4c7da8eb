#!/usr/bin/env bash
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2n_tests.log
tail -n 6 gpurun_out/r2n_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2n_smoke.log 2>&1; tail -n 5 gpurun_out/r2n_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n_bench_2gpu.json 2> gpurun_out/r2n_bench_2gpu.err
python - <<'PY'
import json
for line in open('gpurun_out/r2n_bench_2gpu.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
PY
tail -n 3 gpurun_out/r2n_bench_2gpu.err
