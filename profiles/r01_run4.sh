set -x
timeout 900 python -m pytest tests/test_gpu_full_size.py -x -q -s 2>&1 | tail -25
timeout 600 python profiles/gpu_baseline.py > gpurun_out/r4_gpu_baseline.json 2> gpurun_out/r4_gpu_baseline.err; cat gpurun_out/r4_gpu_baseline.json; tail -3 gpurun_out/r4_gpu_baseline.err
timeout 900 python profiles/extra_configs.py --skip4 > gpurun_out/r4_extra.json 2> gpurun_out/r4_extra.err; cat gpurun_out/r4_extra.json; tail -3 gpurun_out/r4_extra.err
