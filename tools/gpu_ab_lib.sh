#!/bin/bash
# A/B of whole-library builds (tools/_build/libpwpp_b200_<tag>.so, selected through PWPP_LIB): stage times of the KITTI batch and of the
# recorded-scan batch. usage: LIBS="default tag1 tag2" bash tools/gpu_ab_lib.sh
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
for tag in ${LIBS:-default}; do
  if [ "$tag" = default ]; then unset PWPP_LIB; else export PWPP_LIB=$PWD/tools/_build/libpwpp_b200_$tag.so; fi
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --streaming "" > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/ab.json"))
    print(sys.argv[1], "->", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()})
    k = d.get("kitti_scans") or {}
    print("   kitti_scans", round(k.get("ms_per_step", 0), 3), {a: round(b, 3) for a, b in (k.get("stage_ms") or {}).items()}, "dense", round((d.get("dense1m") or {}).get("ms_per_step", 0), 3))
except Exception as e:
    print(sys.argv[1], "ERR", e, open("gpurun_out/ab.err").read()[-1500:])
PY
done
