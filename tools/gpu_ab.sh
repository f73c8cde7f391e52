#!/bin/bash
# A/B of the remaining switches on one box: stage times of the default workload under each setting, + latency probe
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_streams.py tests/test_gpu_reference.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -12
for cfg in ${AB_CFGS:-"PWPP_FRONT=1" "PWPP_FRONT=0"}; do
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
timeout 120 python tools/gpu_latency_probe.py 2>&1 | tail -4
echo "--- latency with the CTA-per-patch kernels (PWPP_FIT_PATCH=1)"
PWPP_FIT_PATCH=1 timeout 120 python tools/gpu_latency_probe.py 2>&1 | tail -4
