#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + per-opcode executed mix from the source page)."""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if kern and kern not in name:
        continue
    print("----", name[:50])
    for w in want:
        if w in idx:
            print(f"   {w:70s} {r[idx[w]]} {units[idx[w]]}")
if kern:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    # first kernel instance only
    hdr = rows[1]
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            data.append(r)
    ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    c, cs = Counter(), Counter()
    for r in data:
        op = r[isrc].strip().split()
        if not op:
            continue
        o = op[1] if op[0].startswith("@") else op[0]
        o = o.split(".")[0]
        c[o] += int(r[ia])
        cs[o] += int(r[ist])
    tot, tots = sum(c.values()), sum(cs.values()) or 1
    print("total warp instructions", tot)
    for o, n in c.most_common(16):
        print(f"   {o:10s} {100 * n / tot:6.2f}% of instr   {100 * cs[o] / tots:6.2f}% of stall samples")
    scols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = Counter()
    for r in data:
        for i in scols:
            try:
                agg[hdr[i]] += int(r[i])
            except ValueError:
                pass
    print("   stalls:", ", ".join(f"{k}={v}" for k, v in agg.most_common(6)))
