set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02s2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s2_pytest_gpu.log
python bench.py > gpurun_out/r02s2_bench.json 2> gpurun_out/r02s2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02s2_bench.err
python tools/exp_round.py lk 64 > gpurun_out/r02s2_exp_lk.log 2>&1; tail -3 gpurun_out/r02s2_exp_lk.log
