#!/usr/bin/env python
"""Registers / spills per kernel from a `make gpu EXTRA_NVFLAGS="-Xptxas -v"` log: python scripts/ptxas_report.py build.log [name filter]"""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for b in re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if flt and flt not in dem:
        continue
    m = re.search(r"Used (\d+) registers", b)
    sp = re.search(r"(\d+) bytes spill stores", b)
    print(f"{m.group(1) if m else '?':>4} regs  spill {sp.group(1) if sp else '?':>4} B  {dem[:150]}")
