set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lk_dense.py -q -x 2>&1 | tail -2
for cfg in "6 3" "5 3" "4 3" "5 0" "5 2" "5 4" "5 5" "6 2"; do set -- $cfg
LKF_DN_MINB=$1 LKF_DN_OWN=$2 timeout 600 python tools/bench_configs.py c4 2> gpurun_out/r02s14_c4.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())['c4']; print('MINB=$1 OWN=$2', round(d['gpu_ms_dense_entry'],4), round(d['gpu_ms_one_iteration_variant'],4))"
done
LKF_DN_MINB=5 LKF_DN_OWN=5 timeout 600 python -m pytest tests/test_gpu_lk_dense.py -q -x 2>&1 | tail -2
