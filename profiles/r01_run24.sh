set -x
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -3
timeout 600 python profiles/backward_timing.py --eager-img 0 > gpurun_out/s24_bwd.json 2> gpurun_out/s24_bwd.err; cat gpurun_out/s24_bwd.json; tail -5 gpurun_out/s24_bwd.err
TPR_PT_DEPTH=96 timeout 300 python profiles/phase_timing.py fp32 2>&1 | tail -21 | tee gpurun_out/s24_phase96.txt
timeout 300 python profiles/phase_timing.py fp32 2>&1 | tail -21 | tee gpurun_out/s24_phase48.txt
