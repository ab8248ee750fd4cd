set -x
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -s 2>&1 | tail -16
timeout 600 python profiles/backward_timing.py --eager-img 0 > gpurun_out/s23_bwd.json 2> gpurun_out/s23_bwd.err; cat gpurun_out/s23_bwd.json; tail -5 gpurun_out/s23_bwd.err
timeout 600 python profiles/backward_timing.py --eager-img 0 --no-keep > gpurun_out/s23_bwd_nokeep.json 2> gpurun_out/s23_bwd.err; cat gpurun_out/s23_bwd_nokeep.json; tail -5 gpurun_out/s23_bwd.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s23_bench.json 2> gpurun_out/s23_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s23_bench.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['train_step'])"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
