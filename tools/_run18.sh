set -u
timeout 300 python tools/exp_round.py lk 64 2>&1 | tail -1
LKFLOW_B200_SO=$PWD/.exp/lib_lkw2.so timeout 300 python tools/exp_round.py lk 64 2>&1 | tail -1
timeout 300 python tools/_time_me.py 2>&1 | tail -1
LKFLOW_B200_SO=$PWD/.exp/lib_me10.so timeout 300 python tools/_time_me.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_gftt.py -q -x 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_vp_frame -c 60 --csv --log-file gpurun_out/r02_vp_c1_tp500.csv python tools/time_vp_c1.py 40 > gpurun_out/r02_vp_c1.log 2>&1; tail -1 gpurun_out/r02_vp_c1.log
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_vp_c1_tp500.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]; h=rows[hi]
v=[float(r[h.index('Metric Value')].replace(',',''))/1000 for r in rows[hi+1:] if len(r)>h.index('Metric Value')]
print('k_vp_frame TP_NUM=500: launches', len(v), 'mean us', sum(v)/len(v), 'max us', max(v), 'last 10 mean', sum(v[-10:])/10)
PY
