N=${1:-8}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/s31_bench_${N}gpu.json 2> gpurun_out/s31_bench_${N}gpu.err
echo "rc=$? bytes=$(wc -c < gpurun_out/s31_bench_${N}gpu.json)"
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/s31_bench_${N}gpu.err | tail -25 | cut -c1-250
cut -c1-300 gpurun_out/s31_bench_${N}gpu.json
