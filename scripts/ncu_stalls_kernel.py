import collections, csv, io, subprocess, sys
rep=sys.argv[1]; which=int(sys.argv[2]); ntop=int(sys.argv[3])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
start = hdr[which]
end = hdr[which+1] - 1 if len(hdr) > which+1 else len(rows)
hh = rows[start]
col = {n: i for i, n in enumerate(hh)}
data = [x for x in rows[start + 1:end] if len(x) > 10]
tot = sum(int(x[col["# Samples"]] or 0) for x in data)
stalls = [n for n in hh if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.Counter()
for x in data:
    for s in stalls:
        agg[s] += int(x[col[s]] or 0)
print("total samples", tot, "instructions", len(data))
for s, v in agg.most_common(12):
    print(f"  {s:28s} {v:8d} {100 * v / max(tot, 1):5.1f}%")
ex = sum(int(x[col["Instructions Executed"]] or 0) for x in data)
print("warp inst", ex)
top = sorted(data, key=lambda r:-int(r[col["# Samples"]] or 0))[:ntop]
for r in top:
    n=int(r[col["# Samples"]]); main=max(stalls,key=lambda s:int(r[col[s]] or 0))
    print(f"{n:6d} {100*n/tot:4.1f}% {main:18s} ex {r[col['Instructions Executed']]:>9s} {r[col['Source']][:90]}")
# opcode histogram by executed count
op=collections.Counter(); ops=collections.Counter()
for r in data:
    s=r[col['Source']].split()
    s=[t for t in s if not t.startswith('@')]
    if not s: continue
    o=s[0].split('.')[0]
    op[o]+=int(r[col["Instructions Executed"]] or 0); ops[o]+=int(r[col["# Samples"]] or 0)
print("opcode: executed share, sample share")
for o,v in op.most_common(30): print(f"  {o:12s} {100*v/ex:5.1f}% {100*ops[o]/tot:5.1f}%")
