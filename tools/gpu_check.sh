#!/bin/bash
# GPU check of HEAD: full -m gpu suite, default bench with all records, latency probe. Every leg under its own timeout.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -12
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_full.json"))
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["gpu_launches"])
    print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_gbs_per_rank", "numa_node", "copying_getters_ms")})
    print("streaming", d.get("streaming"))
    print("dense", {k: v for k, v in (d.get("dense1m") or {}).items() if k not in ("workload",)})
    print("latency", d.get("latency_us"))
    print("parity", d.get("parity_vs_reference"))
    print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), "whole", round(d["roofline"]["whole_path_frac"], 4), {k: (round(v["ms"], 3), round(v["frac"], 3)) for k, v in d["roofline"]["per_kernel"].items()})
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print("bench parse error", e)
PY
timeout 120 python tools/gpu_latency_probe.py 2>&1 | tail -4
