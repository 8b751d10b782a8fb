set -u
timeout 600 python -m pytest tests/test_gpu_lk_dense.py -q -x 2>&1 | tail -2
for mb in 6 5; do
LKF_DN_MINB=$mb timeout 600 python tools/bench_configs.py c4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())['c4']; print('MINB=$mb', {k:round(v,4) for k,v in d.items() if k.startswith('gpu_ms')})"
done
