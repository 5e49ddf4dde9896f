#!/bin/bash
# round 2, occupancy experiments on the 0.5 degree workload: register caps of k_mobi_column / k_mobi_cell / k_diffuse
# (variant libraries built by scripts/build_variants.py) and the warp-specialised MOBI kernel on a large grid
set -u
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0.3"
V=/root/repo/uvic2.9_b200/variants
$B > $O/f_base.json 2> $O/f_base.err
UVIC_B200_LIB=$V/libuvic_b200_A.so $B > $O/f_A.json 2> $O/f_A.err
UVIC_B200_LIB=$V/libuvic_b200_B.so $B > $O/f_B.json 2> $O/f_B.err
UVIC_B200_MOBI_WS=1 $B > $O/f_ws4.json 2> $O/f_ws4.err
UVIC_B200_MOBI_WS=1 UVIC_B200_MOBI_WS_G=1 $B > $O/f_ws1.json 2> $O/f_ws1.err
for v in A B; do
  UVIC_B200_LIB=$V/libuvic_b200_$v.so timeout 400 python -m pytest tests/test_gpu_mobi.py tests/test_gpu_parity.py -q -x > $O/f_pytest_$v.log 2>&1; tail -2 $O/f_pytest_$v.log
done
python - <<'PY'
import json
for t in ("base", "A", "B", "ws4", "ws1"):
    try:
        d = json.loads(open(f"gpurun_out/f_{t}.json").read().strip().splitlines()[-1])
        k = {x["kernel"]: round(x["ms_total"] / d["steps"], 3) for x in d.get("kernels", [])[:8]}
        print(t, d["ms_per_step"], k)
    except Exception as e:
        print(t, "failed", e)
PY
