set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lk_dense.py tests/test_gpu_gftt.py tests/test_gpu_golden.py tests/test_gpu_random.py -q -x > gpurun_out/r02s8_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02s8_pytest.log
for mb in 4 5 6; do
LKF_DN_MINB=$mb timeout 600 python tools/bench_configs.py c4 2> gpurun_out/r02s8_c4.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())['c4']; print('MINB=$mb', {k:round(v,4) for k,v in d.items() if k.startswith('gpu_ms')})"
done
LKF_DN_MINB=6 timeout 300 python -m pytest tests/test_gpu_lk_dense.py -q -x 2>&1 | tail -2
timeout 600 python tools/bench_configs.py c3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())['c3']; print('C3', d['gpu_ms_per_pair'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_sort|k_top|k_nms|k_min_eigen" -c 12 --csv --log-file gpurun_out/r02s8_gftt_launches.csv python tools/prof_one.py gftt 8 > /dev/null 2>&1
grep -E "k_sort|k_top|k_nms|k_min" gpurun_out/r02s8_gftt_launches.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -8
