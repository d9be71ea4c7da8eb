#!/usr/bin/env bash
# round-2 GPU call B: packed warp kernels, fp64 vertex backward, weight-stream experiments
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_warp_gpu.py tests/test_vertex_gpu.py tests/test_mlp_gpu.py -x -q -m gpu > gpurun_out/r2b_tests1.log 2>&1
echo "tests1 exit $?" >> gpurun_out/r2b_tests1.log
timeout 900 python -m pytest tests/test_bench_config_gpu.py tests/test_render_gpu.py -q -m gpu > gpurun_out/r2b_tests2.log 2>&1
echo "tests2 exit $?" >> gpurun_out/r2b_tests2.log
for dbg in 0 1 3 5 9; do
  OCCNERF_MLP_DEBUG=$dbg timeout 120 python tools/mlp_weight_exp.py >> gpurun_out/r2b_weight_exp.jsonl 2>> gpurun_out/r2b_weight_exp.err
done
timeout 300 python tools/diag_point_dist.py > gpurun_out/r2b_diag_pd.log 2>&1
timeout 600 python bench.py --engine tf32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_tf32.json 2> gpurun_out/r2b_bench_tf32.err
tail -n 3 gpurun_out/r2b_tests1.log gpurun_out/r2b_tests2.log
cat gpurun_out/r2b_weight_exp.jsonl
