#!/bin/bash
# One bounded GPU session: parity suite, default bench, A/B of the kernel-variant switches, ncu launch list, dense bench.
# Every leg has its own timeout and writes under gpurun_out/; later legs still run when an earlier one fails.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/gpu.txt 2>&1
t0=$SECONDS
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$((SECONDS-t0))s" | tee -a gpurun_out/legs.txt
t0=$SECONDS
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? t=$((SECONDS-t0))s" | tee -a gpurun_out/legs.txt
t0=$SECONDS
timeout 150 python tools/gpu_ab.py 1024 5 > gpurun_out/ab.log 2>&1; echo "ab rc=$? t=$((SECONDS-t0))s" | tee -a gpurun_out/legs.txt
t0=$SECONDS
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 33 -c 11 --csv --log-file gpurun_out/launches_1024frames.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu-launches rc=$? t=$((SECONDS-t0))s" | tee -a gpurun_out/legs.txt
t0=$SECONDS
timeout 120 python bench.py --sensor dense1m --frames-per-gpu 32 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_dense1m.json 2> gpurun_out/bench_dense1m.err; echo "dense rc=$? t=$((SECONDS-t0))s" | tee -a gpurun_out/legs.txt
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json | cut -c1-600
