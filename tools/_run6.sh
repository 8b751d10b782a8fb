set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02s6_pytest.log
timeout 600 python tools/bench_configs.py c3 c4 > gpurun_out/r02s6_cfg.json 2> gpurun_out/r02s6_cfg.err; echo "cfg rc=$?"; cat gpurun_out/r02s6_cfg.json; tail -3 gpurun_out/r02s6_cfg.err
timeout 600 python bench.py --no-cpu > gpurun_out/r02s6_bench.json 2> gpurun_out/r02s6_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s6_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'])
for k,v in d['stages'].items(): print(k, round(v['ms'],4))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_sort|k_top|k_nms|k_min_eigen|k_masked" -c 40 --csv --log-file gpurun_out/r02s6_gftt_launches.csv python tools/prof_one.py gftt 8 > /dev/null 2>&1
grep -E "k_sort|k_top|k_nms|k_min|k_masked" gpurun_out/r02s6_gftt_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
