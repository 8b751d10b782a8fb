mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print({k:d[k] for k in ['value','n_gpus','ms_per_step']}, d['e2e']['value'], d['h2d_pinned'])
print(d['c5'])
PY
