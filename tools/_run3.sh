set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lk_dense.py tests/test_gpu_lk.py -q -x > gpurun_out/r02s3_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02s3_pytest.log
timeout 600 python tools/bench_configs.py c4 > gpurun_out/r02s3_c4.json 2> gpurun_out/r02s3_c4.err; echo "c4 rc=$?"; cat gpurun_out/r02s3_c4.json; tail -3 gpurun_out/r02s3_c4.err
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_lk_track_t --launch-skip 2 -c 1 -o gpurun_out/r02e_lk python tools/prof_one.py lk 9 > gpurun_out/r02e_lk.log 2>&1; echo "lk rc=$?"
ncu -i gpurun_out/r02e_lk.ncu-rep --page raw --csv > gpurun_out/r02e_lk_raw.csv 2>/dev/null
ncu -i gpurun_out/r02e_lk.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02e_lk_src.csv 2>/dev/null
$NCU -k regex:k_lk_dense -c 1 -o gpurun_out/r02e_dn python tools/prof_one.py dense 2 > gpurun_out/r02e_dn.log 2>&1; echo "dn rc=$?"
ncu -i gpurun_out/r02e_dn.ncu-rep --page raw --csv > gpurun_out/r02e_dn_raw.csv 2>/dev/null
ncu -i gpurun_out/r02e_dn.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02e_dn_src.csv 2>/dev/null
