#!/bin/bash
# Short confirmation of the compiled-in defaults: GPU parity suite, default bench, dense bench, launch list if time allows.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
leg() { echo "$1 rc=$2 t=$((SECONDS-t0))s" | tee -a gpurun_out/legs_confirm.txt; }
: > gpurun_out/legs_confirm.txt
t0=$SECONDS; timeout 90 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; leg pytest $?
t0=$SECONDS; timeout 60 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; leg bench $?
t0=$SECONDS; timeout 30 python bench.py --sensor dense1m --frames-per-gpu 32 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_dense1m_final.json 2> gpurun_out/bench_dense1m_final.err; leg dense $?
if [ $SECONDS -lt ${LAUNCH_LIST_BEFORE:-72} ]; then
t0=$SECONDS; timeout 65 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 33 -c 11 --csv --log-file gpurun_out/launches_1024frames_final.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_final.log 2>&1; leg ncu-launches $?
fi
cat gpurun_out/legs_confirm.txt; tail -3 gpurun_out/pytest_gpu_final.log; cut -c1-260 gpurun_out/bench_n1_final.json
