#!/bin/bash
# ncu evidence of the compiled-in defaults (no PWPP_* switches set), every leg under its own timeout:
#   (1) launch list of one 1024-frame KITTI step and of one 32-frame dense step (gpu__time_duration.sum)
#   (2) ncu --set full of one 128-frame KITTI step (254 MB of input: larger than the 126 MB L2) and one 8-frame dense step
#   (3) raw CSV export, per-kernel summary + DRAM bytes per launch (tools/ncu_summary.py), line-level stall / instruction tables
# Everything lands in gpurun_out/r02_*; nothing printed by a run under ncu is a bench value.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
unset $(env | grep -o '^PWPP_[A-Z0-9_]*' | grep -v PWPP_LIB)
B="python bench.py --no-e2e --no-cpu-baseline --no-extras"
# one step = 9 (cluster front end) or 11 launches; capture the last 2 steps' worth and keep the final one when summarising
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r02_launches_1024frames.csv $B --steps 1 --warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1; echo "launches-kitti rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r02_launches_dense32.csv $B --sensor dense1m --frames-per-gpu 32 --steps 1 --warmup 3 > gpurun_out/r02_ncu_launches_dense.log 2>&1; echo "launches-dense rc=$?"
python - <<'PY'
import csv, glob
for p in sorted(glob.glob("gpurun_out/r02_launches_*.csv")):
    rows = [r for r in csv.reader(open(p)) if len(r) > 5 and r[0].isdigit()]
    names = [r[4] for r in rows]
    # the last step = the launches after the last occurrence of the first kernel of a step
    first = names[0].split("(")[0] if names else ""
    starts = [i for i, n in enumerate(names) if n.split("(")[0] == first]
    last = rows[starts[-1]:] if starts else []
    tot = sum(float(r[-1]) for r in last)
    print(p, "launches/step", len(last), "sum_us", round(tot / 1e3 if tot > 1e5 else tot, 1))
    for r in last:
        print("   ", r[4].split("(")[0][:60], r[-2], r[-1])
PY
N=$(python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_1024frames.csv")) if len(r) > 5 and r[0].isdigit()]
names = [r[4].split("(")[0] for r in rows]
starts = [i for i, n in enumerate(names) if n == names[0]]
print(len(rows) - starts[-1], starts[-1])
PY
)
PER=${N% *}; SKIP=${N#* }
echo "per-step launches $PER, skip $SKIP"
timeout 420 ncu --set full --clock-control none --import-source on -k regex:k_ -s $SKIP -c $PER -f -o gpurun_out/r02_full128 $B --frames-per-gpu 128 --steps 1 --warmup 3 > gpurun_out/r02_ncu_full.log 2>&1; echo "full-kitti rc=$?"
ND=$(python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_dense32.csv")) if len(r) > 5 and r[0].isdigit()]
names = [r[4].split("(")[0] for r in rows]
starts = [i for i, n in enumerate(names) if n == names[0]]
print(len(rows) - starts[-1], starts[-1])
PY
)
PERD=${ND% *}; SKIPD=${ND#* }
timeout 420 ncu --set full --clock-control none --import-source on -k regex:k_ -s $SKIPD -c $PERD -f -o gpurun_out/r02_full_dense8 $B --sensor dense1m --frames-per-gpu 8 --steps 1 --warmup 3 > gpurun_out/r02_ncu_full_dense.log 2>&1; echo "full-dense rc=$?"
for r in r02_full128 r02_full_dense8; do
  [ -f gpurun_out/$r.ncu-rep ] || continue
  timeout 90 ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2> /dev/null
  ls -la gpurun_out/$r.ncu-rep
done
PTS=$(python -c "
import json,sys
try: print(int(json.load(open('gpurun_out/r02_pts128.json'))['points']))
except Exception: print(0)")
[ -f gpurun_out/r02_full128_raw.csv ] && python tools/ncu_summary.py gpurun_out/r02_full128_raw.csv gpurun_out/r02_ncu_full_128frames_summary.csv --traffic gpurun_out/r02_traffic.json --frames 128 --tag "profiles/r02_ncu_full_128frames_summary.csv (ncu --set full, 128 KITTI-shaped frames per launch, compiled-in defaults)"
[ -f gpurun_out/r02_full_dense8_raw.csv ] && python tools/ncu_summary.py gpurun_out/r02_full_dense8_raw.csv gpurun_out/r02_ncu_full_dense8_summary.csv --traffic gpurun_out/r02_traffic_dense.json --frames 8 --tag "profiles/r02_ncu_full_dense8_summary.csv (ncu --set full, 8 dense 1.19 M-point frames per launch, compiled-in defaults)"
if [ -f gpurun_out/r02_full128.ncu-rep ]; then
  for k in k_front_cluster k_bin_hist k_scatter k_fit_cta k_fit_warp k_fit_resident k_emit; do
    timeout 40 python tools/ncu_lines.py gpurun_out/r02_full128.ncu-rep $k 25 0 > gpurun_out/r02_lines_$k.txt 2>&1
    timeout 40 python tools/ncu_inst_lines.py gpurun_out/r02_full128.ncu-rep $k 25 0 > gpurun_out/r02_inst_$k.txt 2>&1
  done
fi
# the merge back is limited to 64 MiB in total
du -sh gpurun_out | tail -1
for r in r02_full128 r02_full_dense8; do [ -f gpurun_out/$r.ncu-rep ] && [ $(stat -c %s gpurun_out/$r.ncu-rep) -gt 24000000 ] && rm -f gpurun_out/$r.ncu-rep; done
rm -f gpurun_out/r02_full128_raw.csv.tmp
ls gpurun_out | grep r02_ | head -40
