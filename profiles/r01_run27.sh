set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/s27_bench_fp32.json 2> gpurun_out/s27_bench_fp32.err; cat gpurun_out/s27_bench_fp32.json; tail -3 gpurun_out/s27_bench_fp32.err
timeout 600 python bench.py --mode bf16 --no-cpu-baseline > gpurun_out/s27_bench_bf16.json 2> gpurun_out/s27_bench_bf16.err; python -c "
import json; d=json.load(open('gpurun_out/s27_bench_bf16.json')); print('bf16', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['train_step']['ms_per_step'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s27_bench_ref.json 2>&1; cut -c1-300 gpurun_out/s27_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s27_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/s27_ncu_b.log 2>&1
python profiles/launch_summary.py gpurun_out/s27_launches.csv > gpurun_out/s27_launches_summary.txt; head -8 gpurun_out/s27_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ws_kernel -c 1 -s 3 -o gpurun_out/s27_render_ws_fp32 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-step > /dev/null 2> gpurun_out/s27_ncu2.err; tail -2 gpurun_out/s27_ncu2.err
timeout 900 python profiles/extra_configs.py --n-img 8 > gpurun_out/s27_extra.json 2> gpurun_out/s27_extra.err; cat gpurun_out/s27_extra.json; tail -3 gpurun_out/s27_extra.err
