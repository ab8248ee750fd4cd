set -x
timeout 600 python -m pytest tests/test_gpu_host_path.py -x -q 2>&1 | tail -15
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_fp32.json 2> gpurun_out/r2_bench_fp32.err; cat gpurun_out/r2_bench_fp32.json; tail -3 gpurun_out/r2_bench_fp32.err
timeout 600 python profiles/phase_timing.py fp32 bf16 > gpurun_out/r2_phase.txt 2>&1; cat gpurun_out/r2_phase.txt
timeout 900 python profiles/extra_configs.py > gpurun_out/r2_extra.json 2> gpurun_out/r2_extra.err; cat gpurun_out/r2_extra.json; tail -3 gpurun_out/r2_extra.err
