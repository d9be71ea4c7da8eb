#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_train_step_gpu.py -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2i_tests.log
tail -n 12 gpurun_out/r2i_tests.log
