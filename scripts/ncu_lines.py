#!/usr/bin/env python
"""Per-source-line stall samples of a .ncu-rep (needs -lineinfo + --import-source on):  python scripts/ncu_lines.py rep [n_top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""
lines = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0] not in ("", "Function Name") and len(r) > 8 and r[0].isdigit():
        i = hdr.index("# Samples")
        stall_cols = [(j, n) for j, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
        num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
        n = num(r[i])
        top = sorted(((num(r[j]), nm) for j, nm in stall_cols), reverse=True)[:2]
        lines.append((n, cur_file, r[0], r[1].strip()[:110], top, r[hdr.index("Instructions Executed")]))
tot = sum(l[0] for l in lines)
print("total samples", tot)
for n, f, ln, srcl, top, ex in sorted(lines, reverse=True)[:ntop]:
    print(f"{n:6d} {100*n/max(tot,1):5.1f}% {f}:{ln:>4s} {top[0][1][6:]:>10s}/{top[1][1][6:]:<10s} | {srcl}")
