#!/usr/bin/env bash
# 2-GPU: train-step tests on one GPU first, then the data-parallel bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_train_step_gpu.py -x -q -m gpu > gpurun_out/r2j_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2j_tests.log
tail -n 6 gpurun_out/r2j_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err
tail -n 5 gpurun_out/r2j_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_2gpu.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
PY
