#!/usr/bin/env bash
# round-2 GPU call A: tf32 engine + bench-config parity + diagnostics
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_mlp_gpu.py tests/test_bench_config_gpu.py -x -q -m gpu > gpurun_out/r2a_tests1.log 2>&1
echo "tests1 exit $?" >> gpurun_out/r2a_tests1.log
timeout 600 python -m pytest tests/test_render_gpu.py -q -m gpu > gpurun_out/r2a_tests2.log 2>&1
echo "tests2 exit $?" >> gpurun_out/r2a_tests2.log
timeout 300 python tools/bench_mlp.py > gpurun_out/r2a_bench_mlp.json 2> gpurun_out/r2a_bench_mlp.err
timeout 300 python tools/mlp_stalls.py > gpurun_out/r2a_mlp_stalls.json 2> gpurun_out/r2a_mlp_stalls.err
timeout 300 python tools/diag_point_dist.py > gpurun_out/r2a_diag_pd.log 2>&1
timeout 600 python tools/hashgrid_vs_ref.py > gpurun_out/r2a_hashgrid_vs_ref.log 2>&1
timeout 600 python bench.py --engine tf32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_tf32.json 2> gpurun_out/r2a_bench_tf32.err
tail -3 gpurun_out/r2a_tests1.log gpurun_out/r2a_tests2.log
cat gpurun_out/r2a_bench_mlp.json
