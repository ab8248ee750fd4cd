set -x
timeout 600 python -m pytest tests/test_gpu_run_model_ws.py tests/test_gpu_parity.py -x -q 2>&1 | tail -15
timeout 900 python profiles/extra_configs.py --skip4 > gpurun_out/s6_extra.json 2> gpurun_out/s6_extra.err; cat gpurun_out/s6_extra.json; tail -3 gpurun_out/s6_extra.err
