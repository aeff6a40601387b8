#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` logs under build/: kernel, registers, spills, smem."""
import glob, re, subprocess, sys
rows = []
for f in sorted(glob.glob('build/*.ptxas.log')):
    cur = None
    for line in open(f):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1); continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur: spill = (int(m.group(2)), int(m.group(3))); rows.append([cur, None, spill, 0]); continue
        m = re.search(r"Used (\d+) registers(?:.*?(\d+) bytes smem)?", line)
        if m and cur and rows and rows[-1][0] == cur:
            rows[-1][1] = int(m.group(1)); rows[-1][3] = int(m.group(2) or 0)
names = subprocess.run(['c++filt'] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for n, r in zip(names, rows):
    n = re.sub(r'\(.*', '', n)
    print(f"{r[1]:4d} regs  spill {r[2][0]:4d}/{r[2][1]:4d}  smem {r[3]:6d}  {n}")
