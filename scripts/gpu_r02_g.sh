#!/bin/bash
# round 2: k_diffuse with a 32 i x 4 level CTA tile (L1 reuse of the rows a level shares with its neighbours) against the
# linear cell order, with 3 instead of 4 CTAs per SM; k_mobi_cell at 64 and 160 registers; parity of the tiled kernel
set -u
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0.3"
V=/root/repo/uvic2.9_b200/variants
$B > $O/g_tiled.json 2> $O/g_tiled.err
UVIC_B200_UPD_TILE=0 $B > $O/g_linear.json 2> $O/g_linear.err
UVIC_B200_LIB=$V/libuvic_b200_C.so $B > $O/g_C.json 2> $O/g_C.err
UVIC_B200_LIB=$V/libuvic_b200_D.so $B > $O/g_D.json 2> $O/g_D.err
UVIC_B200_LIB=$V/libuvic_b200_D.so UVIC_B200_UPD_TILE=0 $B > $O/g_Dlin.json 2> $O/g_Dlin.err
UVIC_B200_UPD_TILE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchsize.py tests/test_gpu_refpin.py -q -x > $O/g_pytest_tiled.log 2>&1; tail -3 $O/g_pytest_tiled.log
python - <<'PY'
import json
for t in ("tiled", "linear", "C", "D", "Dlin"):
    try:
        d = json.loads(open(f"gpurun_out/g_{t}.json").read().strip().splitlines()[-1])
        k = {x["kernel"]: round(x["ms_total"] / d["steps"], 3) for x in d.get("kernels", [])[:8]}
        print(t, d["ms_per_step"], k)
    except Exception as e:
        print(t, "failed", e)
PY
