#!/usr/bin/env python
"""Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in .ncu-rep files
-> profiles/<tag>_ncu_traffic.json, which bench.py quotes as roofline.traffic, and the per-kernel utilisation figures
-> profiles/<tag>_ncu_kernels.json (FP64 pipe, issue slots, registers, FP64 thread instructions per ocean cell).

    NCU_TAG=r02 python scripts/ncu_traffic.py WORKLOAD=report.ncu-rep[:OCEAN_CELLS] [WORKLOAD=report2.ncu-rep ...]
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = os.environ.get("NCU_TAG", "r02")
OUT = os.path.join(ROOT, "profiles", f"{TAG}_ncu_traffic.json")
OUT2 = os.path.join(ROOT, "profiles", f"{TAG}_ncu_kernels.json")
EXTRA = {"gpu__time_duration.sum": "duration", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
         "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_pct",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "launch__registers_per_thread": "registers",
         "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "smsp__inst_executed.sum": "warp_instructions",
         "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
         "smsp__sass_thread_inst_executed_op_fp64_pred_on.sum": "fp64_thread_instr",
         "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "dfma_thread_instr",
         "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "dadd_thread_instr",
         "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "dmul_thread_instr"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
NAMES = {"k_update<2>": "k_diffuse", "k_update<3>": "k_fct_apply", "k_update<1>": "k_update", "k_update<0>": "k_update"}

out = json.load(open(OUT)) if os.path.exists(OUT) else {}
out2 = json.load(open(OUT2)) if os.path.exists(OUT2) else {}
for arg in sys.argv[1:]:
    wl, rep = arg.split("=", 1)
    ocean = None
    if ":" in rep:
        rep, ocean = rep.rsplit(":", 1)
        ocean = float(ocean)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        short = re.sub(r"^void ", "", name).split("(")[0]
        short = NAMES.get(short, re.sub(r"<.*>", "", short))
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[m]]) * SCALE[units[ix[m]]]
        out.setdefault(wl, {})[short] = int(tot)
        rec = {"dram_bytes": int(tot)}
        for m, key in EXTRA.items():
            if m in ix:
                val = float(r[ix[m]])
                if key == "duration":
                    rec["duration_us"] = round(val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[ix[m]], 1.0), 2)
                else:
                    rec[key] = round(val, 2)
        if ocean and rec.get("fp64_thread_instr") and short.startswith("k_mobi"):
            rec["fp64_thread_instr_per_ocean_cell"] = round(rec["fp64_thread_instr"] / ocean, 1)
        out2.setdefault(wl, {})[short] = rec
json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
json.dump(out2, open(OUT2, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
