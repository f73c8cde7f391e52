#!/usr/bin/env python
"""Source lines of one kernel in an ncu report ranked by EXECUTED warp instructions (tools/ncu_lines.py ranks by stall samples).
usage: tools/ncu_inst_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP_N] [LAUNCH_SKIP]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv", "-k", "regex:" + rx, "--launch-count", "1", "--launch-skip", skip], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": print(r[1][:100]); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-": lines.append((fname, r))
ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
toti = sum(int(r[ii]) for _, r in lines)
print("warp instructions", toti)
byfile = {}
for f, r in lines: byfile[f] = byfile.get(f, 0) + int(r[ii])
print({k: f"{100*v/toti:.1f}%" for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])})
for f, r in sorted(lines, key=lambda fr: -int(fr[1][ii]))[:top]:
    print(f"{100*int(r[ii])/toti:5.1f}% {f}:{r[0]:>4} {r[1].strip()[:120]}")
