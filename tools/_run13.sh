set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_lk_dense -c 1 -o gpurun_out/r02h_dn python tools/prof_one.py dense 2 > gpurun_out/r02h_dn.log 2>&1; echo "dn rc=$?"
ncu -i gpurun_out/r02h_dn.ncu-rep --page raw --csv > gpurun_out/r02h_dn_raw.csv 2>/dev/null
ncu -i gpurun_out/r02h_dn.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02h_dn_src.csv 2>/dev/null
