set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gftt.py tests/test_gpu_golden.py tests/test_gpu_random.py tests/test_gpu_vp.py -q -x 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_sort|k_top|k_nms|k_min_eigen" -c 12 --csv --log-file gpurun_out/r02s20_gftt_launches.csv python tools/prof_one.py gftt 8 > /dev/null 2>&1
grep -E "k_sort|k_top|k_nms|k_min" gpurun_out/r02s20_gftt_launches.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -4
