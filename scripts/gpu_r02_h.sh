#!/bin/bash
# round 2: k_mobi_cell at 64 (new default) / 48 / 40 registers, k_diffuse tile shapes (4 levels, 5 levels, 2 levels x 2 rows),
# k_mobi_column with an L2 prefetch of the next level's inputs
set -u
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0.3"
V=/root/repo/uvic2.9_b200/variants
$B > $O/h_base.json 2> $O/h_base.err
UVIC_B200_UPD_TILE=2 $B > $O/h_tile2.json 2> $O/h_tile2.err
for v in E F G H; do UVIC_B200_LIB=$V/libuvic_b200_$v.so $B > $O/h_$v.json 2> $O/h_$v.err; done
UVIC_B200_UPD_TILE=2 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x > $O/h_pytest_tile2.log 2>&1; tail -2 $O/h_pytest_tile2.log
timeout 300 python -m pytest tests/test_gpu_mobi.py -q -x > $O/h_pytest_mobi.log 2>&1; tail -2 $O/h_pytest_mobi.log
python - <<'PY'
import json
for t in ("base", "tile2", "E", "F", "G", "H"):
    try:
        d = json.loads(open(f"gpurun_out/h_{t}.json").read().strip().splitlines()[-1])
        k = {x["kernel"]: round(x["ms_total"] / d["steps"], 3) for x in d.get("kernels", [])[:6]}
        print(t, round(d["ms_per_step"], 3), k)
    except Exception as e:
        print(t, "failed", e)
PY
