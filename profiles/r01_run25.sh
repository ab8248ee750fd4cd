set -x
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "oracle_and_reference_fixture and (inference_96 or uneven or mid_64 or ffhq_small)" 2>&1 | tail -8
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_host_path.py tests/test_gpu_run_model_ws.py -x -q 2>&1 | tail -6
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_backward.py -x -q -k "reference_fixture and not without" 2>&1 | tail -12
