N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/s32_bench_${N}gpu.json 2> gpurun_out/s32_bench_${N}gpu.err
echo "rc=$? lines=$(wc -l < gpurun_out/s32_bench_${N}gpu.json)"
python -c "
import json; d=json.load(open('gpurun_out/s32_bench_${N}gpu.json')); print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'numa', d['host_numa_node'])"
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/s32_bench_${N}gpu.err | tail -5 | cut -c1-200
timeout 300 python bench.py --steps 20 > gpurun_out/s32_bench_1gpu.json 2> gpurun_out/s32_bench_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/s32_bench_1gpu.json')); print('N=1 value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'numa', d['host_numa_node'])"
lscpu | grep -i "numa\|model name\|^CPU(s)" | head -8
