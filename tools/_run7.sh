set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:"k_sort_select_fast|k_nms_compact4|k_min_eigen_stream" --launch-skip 3 -c 3 -o gpurun_out/r02f_gftt python tools/prof_one.py gftt 8 > gpurun_out/r02f_gftt.log 2>&1; echo "gftt rc=$?"
ncu -i gpurun_out/r02f_gftt.ncu-rep --page raw --csv > gpurun_out/r02f_gftt_raw.csv 2>/dev/null
ncu -i gpurun_out/r02f_gftt.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02f_gftt_src.csv 2>/dev/null
$NCU -k regex:k_lk_dense -c 1 -o gpurun_out/r02f_dn python tools/prof_one.py dense 2 > gpurun_out/r02f_dn.log 2>&1; echo "dn rc=$?"
ncu -i gpurun_out/r02f_dn.ncu-rep --page raw --csv > gpurun_out/r02f_dn_raw.csv 2>/dev/null
ncu -i gpurun_out/r02f_dn.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02f_dn_src.csv 2>/dev/null
$NCU -k regex:k_min_eigen_stream -c 1 -o gpurun_out/r02f_me python tools/prof_one.py min_eigen 8 > gpurun_out/r02f_me.log 2>&1; echo "me rc=$?"
ncu -i gpurun_out/r02f_me.ncu-rep --page raw --csv > gpurun_out/r02f_me_raw.csv 2>/dev/null
ncu -i gpurun_out/r02f_me.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02f_me_src.csv 2>/dev/null
