#!/usr/bin/env python
"""Top source lines of one kernel in an ncu report by warp-stall samples.
usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP_N] [LAUNCH_SKIP]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv", "-k", "regex:" + rx,
                      "--launch-count", "1", "--launch-skip", skip], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": print(r[1][:110]); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        lines.append((fname, r))
if not hdr: sys.exit("no data")
si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for _, r in lines); toti = sum(int(r[ii]) for _, r in lines)
print(f"total samples {tot}, warp instructions {toti}")
agg = {}
for _, r in lines:
    for i in stall: agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("stalls:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for f, r in sorted(lines, key=lambda fr: -int(fr[1][si]))[:top]:
    s = int(r[si]); st = sorted(((int(r[i]), hdr[i][6:]) for i in stall), reverse=True)[:2]
    print(f"{100*s/max(tot,1):5.1f}% inst {100*int(r[ii])/max(toti,1):4.1f}% {f}:{r[0]:>4} [{st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]}] {r[1].strip()[:100]}")
