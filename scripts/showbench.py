#!/usr/bin/env python
"""Print the headline numbers and the per-kernel table of bench.py JSON lines."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f"{f}: value={d['value']:.3f} ms/step={d['ms_per_step']:.3f} e2e={(d.get('e2e') or {}).get('value')} cpu={(d.get('cpu_baseline') or {}).get('value')} launches={d.get('gpu_launches')}")
    r = d.get("roofline") or {}
    print("   roofline:", r.get("kernel"), r.get("achieved"), r.get("frac"))
    for k in d.get("kernels", []):
        print(f"     {k['kernel']:16s} {k['us_per_launch']:10.2f} us x{k['launches']:4d} share={k['share']:.3f} gbs={k['gbs']}")
