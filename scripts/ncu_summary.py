#!/usr/bin/env python
"""Key metrics + stall table of a .ncu-rep:  python scripts/ncu_summary.py gpurun_out/x.ncu-rep [n_top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h = r[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "sm__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum",
        "local_load_requests", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for row in r[2:]:
    print("kernel:", row[h.index("Kernel Name")][:100] if "Kernel Name" in h else "")
    for k in keys:
        if k in h:
            print(f"  {k:70s} {row[h.index(k)]}  {r[1][h.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
if not hdr:
    sys.exit()
start = hdr[0]
end = hdr[1] - 1 if len(hdr) > 1 else len(rows)
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
for s, v in agg.most_common(10):
    print(f"  {s:28s} {v:8d} {100 * v / max(tot, 1):5.1f}%")
ex = sum(int(x[col["Instructions Executed"]] or 0) for x in data)
print("warp instructions executed", ex)
for x in sorted(data, key=lambda x: -int(x[col["# Samples"]] or 0))[:ntop]:
    n = int(x[col["# Samples"]])
    main = max(stalls, key=lambda s: int(x[col[s]] or 0))
    print(f"{n:6d} {100 * n / max(tot, 1):4.1f}% {main:18s} exec {x[col['Instructions Executed']]:>8s}  {x[col['Source']][:100]}")
