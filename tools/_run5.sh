set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lk_dense.py -q -x > gpurun_out/r02s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02s5_pytest.log
for th in 8 4; do
LKF_DN_TH=$th timeout 600 python tools/bench_configs.py c4 2> gpurun_out/r02s5_c4.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())['c4']; print('TH=$th', {k:round(v,4) for k,v in d.items() if k.startswith('gpu_ms')})"
done
LKF_DN_TH=4 timeout 300 python -m pytest tests/test_gpu_lk_dense.py -q -x 2>&1 | tail -2
