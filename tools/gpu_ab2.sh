#!/bin/bash
# r02 A/B on one box: front-end variants (stage times of the default workload), single-frame latency under the remaining switches,
# cycles of one plane solve (tools/solve_bench.cu). Every leg under its own timeout; results in gpurun_out/ab2.log.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
for cfg in ${AB_CFGS:-"PWPP_X=0" "PWPP_FRONT=0" "PWPP_FIT_PATCH=1"}; do
  env $cfg timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/ab.json"))
    print(sys.argv[1], "->", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e, open("gpurun_out/ab.err").read()[-1500:])
PY
done
for cfg in ${LAT_CFGS:-"PWPP_X=0" "PWPP_SMALL_CALL=0" "PWPP_GRAPH=0"}; do
  echo "--- latency probe under $cfg"
  env $cfg timeout 120 python tools/gpu_latency_probe.py 2>&1 | tail -3
done
if [ -x tools/_build/solve_bench ]; then timeout 60 tools/_build/solve_bench; fi
