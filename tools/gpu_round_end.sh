#!/bin/bash
# Last GPU session of a round, every leg bounded: (1) pick the kernel-variant switches on this GPU (tools/gpu_tune.py,
# results checked against the base configuration), (2) GPU parity suite UNDER the chosen switches, (3) bench.py at
# N=1 on the default workload and on the dense workload, (4) ncu launch list of one 1024-frame step, (5) one
# ncu --set full capture of a 128-frame step + per-line stall summaries of the top kernels.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
leg() { echo "$1 rc=$2 t=$((SECONDS-t0))s" | tee -a gpurun_out/legs.txt; }
: > gpurun_out/legs.txt
# (0) the switched-off variants, each in its own process and under its own timeout; only those that reproduce the default
#     configuration's results bit for bit may be timed and chosen by the tuner
: > gpurun_out/variants_ok.txt
if [ "${SKIP_VARIANTS:-0}" != "1" ]; then
  t0=$SECONDS
  for v in "FRONT=1" "PART_ILP=1" "EMIT_SPLIT=8" "SOLVE_CALL=1" "L2_WIDE=1" "X_FIXPOINT=1" "M_RESIDENT=1" "L1_CTA=1" "M_HALF=1" "L2_PLS=1" "L2_PLS=1,L2_MINB=4" "FUSE_SEED=3" "FUSE_SEED=0,L2_MINB=4"; do
    if PWPP_TEST_VARIANTS=1 timeout 120 python -m pytest -m gpu -q -p no:cacheprovider "tests/test_gpu_variants.py::test_variant_equals_default[$v]" > gpurun_out/variant_${v//[=,]/_}.log 2>&1 && grep -q "1 passed" gpurun_out/variant_${v//[=,]/_}.log; then echo "$v" >> gpurun_out/variants_ok.txt; fi
  done
  leg variants $?; cat gpurun_out/variants_ok.txt
fi
export PWPP_TUNE_ALLOW_FILE=gpurun_out/variants_ok.txt
t0=$SECONDS; timeout 150 python tools/gpu_tune.py 1024 5 32 > gpurun_out/tune.log 2>&1; leg tune $?
[ -f gpurun_out/chosen.env ] && source gpurun_out/chosen.env
env | grep '^PWPP_' | sort > gpurun_out/chosen_effective.txt
t0=$SECONDS; timeout 150 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; leg pytest $?
t0=$SECONDS; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; leg smoke $?
t0=$SECONDS; timeout 100 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; leg bench $?
t0=$SECONDS; timeout 60 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --streaming 64x16 > gpurun_out/bench_streaming.json 2> gpurun_out/bench_streaming.err; leg streaming $?
t0=$SECONDS; timeout 60 python bench.py --sensor dense1m --frames-per-gpu 32 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_dense1m.json 2> gpurun_out/bench_dense1m.err; leg dense $?
t0=$SECONDS; timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 33 -c 11 --csv --log-file gpurun_out/launches_1024frames.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; leg ncu-launches $?
CMD="python bench.py --frames-per-gpu 128 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
# the capture gets what is left of the session's time budget (DEADLINE seconds after start, default 600), minus the export
left=$(( ${DEADLINE:-600} - SECONDS - 35 ))
t0=$SECONDS
if [ $left -gt 30 ]; then timeout $left ncu --set full --clock-control none --import-source on -k regex:k_ -s 33 -c 11 -f -o gpurun_out/full128 $CMD > gpurun_out/ncu_full.log 2>&1; leg ncu-full $?; else echo "ncu-full skipped (no time left)" | tee -a gpurun_out/legs.txt; fi
if [ -f gpurun_out/full128.ncu-rep ]; then
  t0=$SECONDS
  timeout 60 ncu -i gpurun_out/full128.ncu-rep --page raw --csv > gpurun_out/full128_raw.csv 2> /dev/null
  for k in k_fit_cta k_fit_warp k_scatter k_bin_hist; do timeout 40 python tools/ncu_lines.py gpurun_out/full128.ncu-rep $k 25 0 > gpurun_out/lines_$k.txt 2>&1; done
  timeout 40 python tools/ncu_lines.py gpurun_out/full128.ncu-rep k_fit_cta 25 1 > gpurun_out/lines_k_fit_cta_2nd.txt 2>&1
  timeout 40 python tools/ncu_lines.py gpurun_out/full128.ncu-rep k_fit_warp 25 1 > gpurun_out/lines_k_fit_warp_2nd.txt 2>&1
  leg ncu-export $?
  ls -la gpurun_out/full128.ncu-rep | tee -a gpurun_out/legs.txt
  # the merge back is limited to 64 MiB: the text exports are what matters if the report is too big
  [ $(stat -c %s gpurun_out/full128.ncu-rep) -gt 45000000 ] && rm -f gpurun_out/full128.ncu-rep
fi
cat gpurun_out/legs.txt; cat gpurun_out/chosen_effective.txt; tail -3 gpurun_out/pytest_gpu.log; cut -c1-300 gpurun_out/bench_n1.json
