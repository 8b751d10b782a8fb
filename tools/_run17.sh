mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus $N"
timeout 900 python bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "rc=$?"; tail -2 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print({k:d[k] for k in ['value','n_gpus','ms_per_step']}, 'e2e', d['e2e']['value'], d['h2d_pinned']['GBps_all_ranks'])
print(d['c5'])
PY
