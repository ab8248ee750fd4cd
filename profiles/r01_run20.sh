set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_frames.py -x -q 2>&1 | tail -5
TPR_PT_DEPTH=96 timeout 300 python profiles/phase_timing.py fp32 2>&1 | tail -21 | tee gpurun_out/s20_phase96.txt
timeout 900 python profiles/extra_configs.py --n-img 8 --skip5 > gpurun_out/s20_extra.json 2> gpurun_out/s20_extra.err; cat gpurun_out/s20_extra.json; tail -3 gpurun_out/s20_extra.err
timeout 300 python bench.py --no-cpu-baseline --no-train-step > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s20_bench.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
