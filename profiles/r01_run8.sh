set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
TPR_PT_DEPTH=96 timeout 300 python profiles/phase_timing.py fp32 2>&1 | tail -25 | tee gpurun_out/s8_phase96.txt
timeout 900 python profiles/extra_configs.py --n-img 8 --skip5 > gpurun_out/s8_extra.json 2> gpurun_out/s8_extra.err; cat gpurun_out/s8_extra.json; tail -3 gpurun_out/s8_extra.err
timeout 300 python bench.py > gpurun_out/s8_bench_fp32.json 2> gpurun_out/s8_bench_fp32.err; cat gpurun_out/s8_bench_fp32.json; tail -3 gpurun_out/s8_bench_fp32.err
