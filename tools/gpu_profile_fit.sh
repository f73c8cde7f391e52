#!/bin/bash
# ncu --set full capture (with source-level sampling) of the two CTA fit kernels on a 128-frame step
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
CMD="python bench.py --frames-per-gpu 128 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fit_cta" -s 6 -c 2 -f -o gpurun_out/prof_cta $CMD > gpurun_out/prof_cta.out 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_cta.ncu-rep; tail -3 gpurun_out/prof_cta.out
