#!/bin/bash
# round-2 GPU session: parity suite, bench with the group fit kernel vs the size-classed kernels, launch list, full capture
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
leg() { echo "$1 rc=$2 t=$((SECONDS-t0))s" | tee -a gpurun_out/legs.txt; }
: > gpurun_out/legs.txt
t0=$SECONDS; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; leg smoke $?
t0=$SECONDS; timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; leg pytest $?
t0=$SECONDS; timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_group.json 2> gpurun_out/bench_group.err; leg bench-group $?
t0=$SECONDS; PWPP_FIT_GROUP=0 timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_classes.json 2> gpurun_out/bench_classes.err; leg bench-classes $?
t0=$SECONDS; timeout 60 python bench.py --sensor dense1m --frames-per-gpu 32 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_dense1m.json 2> gpurun_out/bench_dense1m.err; leg dense $?
t0=$SECONDS; timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 30 -c 10 --csv --log-file gpurun_out/launches_1024frames.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; leg ncu-launches $?
if [ "${FULL:-1}" = "1" ]; then
t0=$SECONDS; timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_fit_group -s 9 -c 3 -f -o gpurun_out/full_group \
  python bench.py --frames-per-gpu 128 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; leg ncu-full $?
if [ -f gpurun_out/full_group.ncu-rep ]; then
  timeout 60 ncu -i gpurun_out/full_group.ncu-rep --page raw --csv > gpurun_out/full_group_raw.csv 2> /dev/null
  for i in 0 1 2; do timeout 40 python tools/ncu_lines.py gpurun_out/full_group.ncu-rep k_fit_group 30 $i > gpurun_out/lines_k_fit_group_$i.txt 2>&1; done
  for i in 0 1 2; do timeout 40 python tools/ncu_inst_lines.py gpurun_out/full_group.ncu-rep k_fit_group 30 $i > gpurun_out/inst_k_fit_group_$i.txt 2>&1; done
fi
fi
cat gpurun_out/legs.txt; tail -3 gpurun_out/pytest_gpu.log; cut -c1-200 gpurun_out/bench_group.json
python - <<'PY'
import json
for n in ("bench_group","bench_classes","bench_dense1m"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json")); print(n, round(d["value"]), d["ms_per_step"], {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, d.get("e2e"))
    except Exception as e: print(n, "ERR", e)
PY
