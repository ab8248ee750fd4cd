set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/s21_bench_2gpu.json 2> gpurun_out/s21_bench_2gpu.err; cat gpurun_out/s21_bench_2gpu.json; tail -3 gpurun_out/s21_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 100 --warmup 5 --gather nccl > gpurun_out/s21_bench_2gpu_nccl.json 2> gpurun_out/s21_bench_2gpu_nccl.err; cat gpurun_out/s21_bench_2gpu_nccl.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
