set -u
mkdir -p gpurun_out
python tools/probe_dense_div.py > gpurun_out/r02s2_dense_div.log 2>&1; cat gpurun_out/r02s2_dense_div.log
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_lk_track_t --launch-skip 3 -c 1 -o gpurun_out/r02e_lk python tools/prof_one.py lk 9 > gpurun_out/r02e_lk.log 2>&1; echo "lk rc=$?"
$NCU -k regex:"k_sort_select|k_nms_compact|k_min_eigen" --launch-skip 3 -c 3 -o gpurun_out/r02e_gftt python tools/prof_one.py gftt 8 > gpurun_out/r02e_gftt.log 2>&1; echo "gftt rc=$?"
$NCU -k regex:k_lk_track_small --launch-skip 1 -c 1 -o gpurun_out/r02e_dense python tools/prof_one.py dense 2 > gpurun_out/r02e_dense.log 2>&1; echo "dense rc=$?"
for n in lk gftt dense; do
  ncu -i gpurun_out/r02e_$n.ncu-rep --page raw --csv > gpurun_out/r02e_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02e_$n.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02e_${n}_src.csv 2>/dev/null
done
ls -la gpurun_out | head -30
