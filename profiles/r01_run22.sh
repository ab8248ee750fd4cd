set -x
timeout 600 python -m pytest tests/test_gpu_host_path.py -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/s22_bench_2gpu.json 2> gpurun_out/s22_bench_2gpu.err; cat gpurun_out/s22_bench_2gpu.json; tail -3 gpurun_out/s22_bench_2gpu.err
