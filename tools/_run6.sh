T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2w_bench_8gpu.json 2> gpurun_out/r2w_bench_8gpu.err
tail -3 gpurun_out/r2w_bench_8gpu.err
