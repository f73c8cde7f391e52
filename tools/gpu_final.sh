#!/bin/bash
# Final evidence pass of the round: GPU tests, both bench arms at N=1, launch list + ncu --set full capture.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
bash tools/gpu_check.sh > gpurun_out/check.log 2>&1; grep -E "rc=|passed|failed|stopping" gpurun_out/check.log
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "ours rc=$?"; cat gpurun_out/bench_full.json | cut -c1-600
CMD="python bench.py --frames-per-gpu 128 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 33 -c 11 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches.out 2>&1; echo "launch list rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 33 -c 11 --csv --log-file gpurun_out/launches_1024.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_1024.out 2>&1; echo "launch list 1024 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ -s 33 -c 11 -f -o gpurun_out/prof $CMD > gpurun_out/prof.out 2>&1; echo "full capture rc=$?"
