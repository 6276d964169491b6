import csv, sys, collections
rows = list(csv.reader(open('/tmp/src_sass.csv')))
# find header rows; process first kernel only
hdr_idx = [i for i,r in enumerate(rows) if r and r[0]=="Address"]
print("kernels:", len(hdr_idx))
start = hdr_idx[0]; end = hdr_idx[1]-1 if len(hdr_idx)>1 else len(rows)
h = rows[start]
col = {n:i for i,n in enumerate(h)}
data = rows[start+1:end]
tot = sum(int(r[col["# Samples"]] or 0) for r in data if len(r)>10)
print("total samples", tot, "instructions", len(data))
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.Counter()
for r in data:
    if len(r) < 10: continue
    for s in stalls:
        agg[s] += int(r[col[s]] or 0)
for s,v in agg.most_common(): print(f"{s:28s} {v:8d} {100*v/tot:5.1f}%")
# top instructions by samples
top = sorted([r for r in data if len(r)>10], key=lambda r:-int(r[col["# Samples"]] or 0))[:45]
for r in top:
    n=int(r[col["# Samples"]]); 
    main = max(stalls, key=lambda s:int(r[col[s]] or 0))
    print(f"{n:6d} {100*n/tot:4.1f}% {main:18s} {r[col['Source']][:110]}")
