python -m pytest tests/test_mlp_gpu.py -x -q -m gpu > gpurun_out/r2p_test_mlp.log 2>&1; tail -3 gpurun_out/r2p_test_mlp.log
python tools/bench_mlp.py > gpurun_out/r2p_bench_mlp.json 2> gpurun_out/r2p_bench_mlp.err
python tools/mlp_trace.py > gpurun_out/r2p_mlp_trace.txt 2>&1
python tools/mlp_trace_w.py > gpurun_out/r2p_trace_w.txt 2>&1
