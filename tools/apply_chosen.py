#!/usr/bin/env python
"""Bakes gpurun_out/chosen.env (written on the GPU box by tools/gpu_tune.py) into csrc/pwpp_tuning.h.
usage: python tools/apply_chosen.py [path/to/chosen.env]"""
import os, re, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
env = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "gpurun_out", "chosen.env")
hdr = os.path.join(REPO, "patchwork-plusplus_b200", "csrc", "pwpp_tuning.h")
src = open(hdr).read()
for line in open(env):
    m = re.match(r"export (PWPP_[A-Z0-9_]+)=(\d+)\s*$", line)
    if not m:
        continue
    name, val = m.group(1) + "_DEFAULT", m.group(2)
    new, n = re.subn(rf"(#define {name} )\d+", rf"\g<1>{val}", src)
    if n != 1:
        sys.exit(f"{name} is not a known switch in pwpp_tuning.h")
    src = new
    print(f"{name} = {val}")
open(hdr, "w").write(src)
