set -x
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -3
timeout 600 python profiles/backward_timing.py --eager-img 0 > gpurun_out/s26_bwd.json 2> gpurun_out/s26_bwd.err; cat gpurun_out/s26_bwd.json; tail -5 gpurun_out/s26_bwd.err
timeout 600 python profiles/backward_timing.py --eager-img 0 --no-keep-features > gpurun_out/s26_bwd_nofeat.json 2> gpurun_out/s26_bwd.err; cat gpurun_out/s26_bwd_nofeat.json; tail -5 gpurun_out/s26_bwd.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s26_bench.json 2> gpurun_out/s26_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s26_bench.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['train_step']['ms_per_step'], d['roofline']['frac_of_l2_gather'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s26_bwd_launches.csv python profiles/backward_timing.py --eager-img 0 --reps 1 > /dev/null 2> gpurun_out/s26_ncu.err; tail -3 gpurun_out/s26_ncu.err
python profiles/launch_summary.py gpurun_out/s26_bwd_launches.csv > gpurun_out/s26_launch_summary.txt; head -8 gpurun_out/s26_launch_summary.txt
