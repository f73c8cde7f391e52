#!/usr/bin/env python
"""SASS instruction count per source line of one kernel (static code size attribution).
usage: tools/sass_lines.py LIB.so KERNEL_SUBSTRING [TOP_N]"""
import collections, glob, os, re, subprocess, sys, tempfile
lib, target = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
cnt = collections.Counter(); line = None; infn = False
for cubin in glob.glob(d + "/*.cubin"):
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    for l in dis.splitlines():
        if ".section" in l and ".text." in l: infn = target in l
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l): cnt[line] += 1
print("total", sum(cnt.values()))
byfile = collections.Counter()
for (f, ln), c in cnt.items(): byfile[f] += c
print(byfile.most_common())
for (f, ln), c in cnt.most_common(top): print(c, f, ln)
