set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lk_dense.py -q -x > gpurun_out/r02s4_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02s4_pytest.log
timeout 600 python tools/bench_configs.py c4 > gpurun_out/r02s4_c4.json 2> gpurun_out/r02s4_c4.err; echo "c4 rc=$?"; cat gpurun_out/r02s4_c4.json; tail -3 gpurun_out/r02s4_c4.err
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_lk_dense -c 1 -o gpurun_out/r02e_dn2 python tools/prof_one.py dense 2 > gpurun_out/r02e_dn2.log 2>&1; echo "dn rc=$?"
ncu -i gpurun_out/r02e_dn2.ncu-rep --page raw --csv > gpurun_out/r02e_dn2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02e_dn2.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02e_dn2_src.csv 2>/dev/null
