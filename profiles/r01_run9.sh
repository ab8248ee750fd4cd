set -x
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -s 2>&1 | tail -40
