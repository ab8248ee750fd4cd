set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/final_bench_fp32.json 2> gpurun_out/final_bench_fp32.err; wc -l gpurun_out/final_bench_fp32.json; python -c "
import json; d=json.load(open('gpurun_out/final_bench_fp32.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_l2_gather'], d['e2e']['value'], d['train_step']['ms_per_step'], d['cpu_baseline']['cores'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-200
