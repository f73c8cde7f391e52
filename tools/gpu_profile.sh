#!/bin/bash
# ncu evidence: launch list (device time of every kernel of one step) + one --set full capture per kernel.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
FR=${1:-128}
CMD="python bench.py --frames-per-gpu $FR --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
# 3 warm-up steps x 10 kernels are skipped, the timed step's 10 kernels are captured
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 30 -c 10 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches.out 2>&1
echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ -s 30 -c 10 -f -o gpurun_out/prof $CMD > gpurun_out/prof.out 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof.ncu-rep
