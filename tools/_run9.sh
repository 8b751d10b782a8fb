set -u
mkdir -p gpurun_out
for t in 0 1; do LKF_LK_TMA=$t timeout 300 python tools/exp_round.py lk 64 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_lk.py tests/test_gpu_random.py -q -x 2>&1 | tail -3
