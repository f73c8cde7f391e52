#!/bin/bash
# Full-size numbers: both arms of bench.py at N=1 (default workload), plus the ncu launch list of one step.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
echo "== ours"; timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo "== launch list (1 step of 1024 frames)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 30 -c 10 --csv --log-file gpurun_out/launches_1024.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_1024.out 2>&1; echo "rc=$?"
