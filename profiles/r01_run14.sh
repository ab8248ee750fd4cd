set -x
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -s 2>&1 | tail -14
timeout 600 python profiles/backward_timing.py --eager-img 0 --mode bf16 > gpurun_out/s14_bwd_bf16.json 2> gpurun_out/s14_bwd.err; cat gpurun_out/s14_bwd_bf16.json; tail -5 gpurun_out/s14_bwd.err
timeout 600 python profiles/backward_timing.py --eager-img 0 > gpurun_out/s14_bwd_fp32.json 2> gpurun_out/s14_bwd.err; cat gpurun_out/s14_bwd_fp32.json; tail -5 gpurun_out/s14_bwd.err
