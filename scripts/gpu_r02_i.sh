#!/bin/bash
# round 2: rows per march chunk (64 default; 96, 128, 180), k_mobi_cell at 40 (default) / 32 registers
set -u
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0.3"
V=/root/repo/uvic2.9_b200/variants
$B > $O/i_base.json 2> $O/i_base.err
for c in 96 128 180; do UVIC_B200_FCT_CHUNK=$c $B > $O/i_c$c.json 2> $O/i_c$c.err; done
UVIC_B200_LIB=$V/libuvic_b200_I.so $B > $O/i_I.json 2> $O/i_I.err
python - <<'PY'
import json
for t in ("base", "c96", "c128", "c180", "I"):
    try:
        d = json.loads(open(f"gpurun_out/i_{t}.json").read().strip().splitlines()[-1])
        k = {x["kernel"]: round(x["ms_total"] / d["steps"], 3) for x in d.get("kernels", [])[:6]}
        print(t, round(d["ms_per_step"], 3), k)
    except Exception as e:
        print(t, "failed", e)
PY
