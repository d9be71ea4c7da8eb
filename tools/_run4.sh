python -m pytest tests/test_deconv_gpu.py tests/test_prologue_gpu.py -x -q -m gpu > gpurun_out/r2t_deconv_tests.log 2>&1; tail -4 gpurun_out/r2t_deconv_tests.log
python tools/decoder_bench.py > gpurun_out/r2t_decoder_bench.json 2> gpurun_out/r2t_decoder_bench.err; tail -3 gpurun_out/r2t_decoder_bench.err
