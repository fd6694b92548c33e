#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` output: registers / spills / stack per kernel (demangled)."""
import re, subprocess, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
cur = None
rows = {}
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1); rows[cur] = {}
        continue
    if cur is None: continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m: rows[cur].update(stack=int(m.group(1)), sst=int(m.group(2)), sld=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m:
        rows[cur]["regs"] = int(m.group(1))
        cur = None
names = list(rows)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
for n, d in zip(names, dem):
    r = rows[n]
    d = re.sub(r"\(.*", "", d).replace("void b200rt::", "")
    print(f"{d:55s} regs {r.get('regs','?'):>3}  stack {r.get('stack',0):>4}  spill st/ld {r.get('sst',0)}/{r.get('sld',0)}")
