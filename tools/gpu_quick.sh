#!/bin/bash
# quick GPU check: bench (stage times), parity subset, optional phase probe with the diagnostic library
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_quick.json"))
    print("BENCH", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()})
except Exception as e:
    print("BENCH ERR", e, open("gpurun_out/bench_quick.err").read()[-2000:])
PY
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
if [ -n "$PROBE" ]; then PWPP_LIB=$PWD/tools/_build/libpwpp_b200_clk.so python tools/gpu_phase_probe.py 256 | tail -1; fi
