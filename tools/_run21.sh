set -u
timeout 300 python tools/_time_me.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_gftt.py tests/test_gpu_golden.py tests/test_gpu_random.py -q -x 2>&1 | tail -3
