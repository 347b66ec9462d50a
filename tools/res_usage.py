#!/usr/bin/env python
"""Registers / stack per kernel of libvllm_ltr_b200.so (cuobjdump -res-usage, no GPU needed).
    python tools/res_usage.py [substring filter]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vllm_ltr_b200", "libvllm_ltr_b200.so")
out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
names, rows = [], []
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        names.append(cur)
        rows.append(tuple(int(x) for x in m.groups()))
        cur = None
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for n, r in zip(dem, rows):
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    if flt in n:
        print(f"{r[0]:4d} regs  stack {r[1]:3d}  local {r[3]:3d}  {n}")
