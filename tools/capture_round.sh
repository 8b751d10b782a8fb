#!/bin/bash
# Round evidence capture on the GPU box (one B200): tests, bench (both arms), ncu launch list of the bench command, ncu --set full
# of exactly one eager step (tools/prof_step.py between cudaProfilerStart/Stop) -> per-kernel DRAM traffic (profiles/<tag>_traffic.json
# through tools/ncu_traffic.py) and counter summaries.  Outputs under gpurun_out/ (scratch); summaries are copied to profiles/.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-c5 --no-dropin --no-configs > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:^k_ -f -o gpurun_out/${TAG}_step \
    python tools/prof_step.py > gpurun_out/${TAG}_ncu_step.log 2>&1; echo "step capture rc=$?"
ncu -i gpurun_out/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_raw.csv
python tools/ncu_traffic.py gpurun_out/${TAG}_step_raw.csv > gpurun_out/${TAG}_traffic.txt 2>&1; cat gpurun_out/${TAG}_traffic.txt | tail -14
cp profiles/r02_traffic.json gpurun_out/${TAG}_traffic.json
