T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $T --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3d_bench_2gpu.json 2> gpurun_out/r3d_bench_2gpu.err
grep "all-reduce phases" gpurun_out/r3d_bench_2gpu.err
