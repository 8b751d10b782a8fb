set -u
mkdir -p gpurun_out
for tl in 32 64 96 128; do
LKF_DN_TAIL=$tl timeout 600 python tools/bench_configs.py c4 2> gpurun_out/r02s12_c4.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())['c4']; print('TAIL=$tl', {k:round(v,4) for k,v in d.items() if k.startswith('gpu_ms')})"
done
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02s12_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02s12_pytest.log
