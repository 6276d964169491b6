#!/usr/bin/env python
"""Warp instructions executed per source line of a .ncu-rep:  python scripts/ncu_inst.py rep [n_top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = "", None, []
num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[0].isdigit():
        lines.append((num(r[hdr.index("Instructions Executed")]), cur_file, r[0], r[1].strip()[:120]))
tot = sum(l[0] for l in lines)
print("warp instructions", tot)
byfile = {}
for n, f, ln, s in lines:
    byfile[f] = byfile.get(f, 0) + n
print({k: f"{100*v/tot:.1f}%" for k, v in byfile.items()})
for n, f, ln, s in sorted(lines, reverse=True)[:ntop]:
    print(f"{n:10d} {100*n/tot:5.1f}% {f}:{ln:>4s} | {s}")
