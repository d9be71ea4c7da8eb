T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 500 $T bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2s_bench_2gpu_switch.json 2> gpurun_out/r2s_bench_2gpu_switch.err
OCCNERF_REDUCER=nccl timeout 500 $T bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2s_bench_2gpu_nccl.json 2> gpurun_out/r2s_bench_2gpu_nccl.err
tail -3 gpurun_out/r2s_bench_2gpu_switch.err
