#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
CMD="python bench.py --frames-per-gpu 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_fit_warp|k_fit_cta|k_fit_resident" -s 12 -c 4 -f -o gpurun_out/prof_fit $CMD > gpurun_out/prof_fit.out 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_fit.ncu-rep
