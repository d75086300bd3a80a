#!/usr/bin/env python3
"""Summarise an .ncu-rep (first kernel): key raw metrics + SASS opcode mix + top stall reasons.
usage: ncu_summary.py file.ncu-rep [warp_iterations]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
iters = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg ", "smsp__cycles_active.avg ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "sm__cycles_elapsed.avg.per_second",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"]
for i, h in enumerate(hdr):
    if any(h == w.strip() or (w.endswith(" ") and h == w.strip()) for w in want):
        print("%-75s %-14s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
byop = collections.Counter(); tot = 0
stalls = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    try:
        n = int(r[ix["Instructions Executed"]])
    except Exception:
        continue
    s = r[ix["Source"]]
    op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
    byop[op] += n; tot += n
    for c in stall_cols:
        try:
            stalls[c] += int(r[ix[c]] or 0)
        except Exception:
            pass
print("total warp instructions", tot, ("per warp-iteration %.1f" % (tot / iters)) if iters else "")
print("opcode mix:", ", ".join("%s %.1f%%" % (op, 100.0 * n / tot) + ((" (%.1f)" % (n / iters)) if iters else "") for op, n in byop.most_common(22)))
st = sum(stalls.values()) or 1
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / st) for k, v in stalls.most_common(8)))
