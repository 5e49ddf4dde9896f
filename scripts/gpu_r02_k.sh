#!/bin/bash
# round 2, closing single-GPU pass after the k_mobi_cell dual allocation: GPU tests (with the launch-geometry test),
# the 100x100x19 lines, their ncu capture, the default bench line
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/k_pytest.log 2>&1; tail -3 $O/k_pytest.log
python bench.py --workload uvic100_mobi37 > $O/k_uvic.json 2> $O/k_uvic.err; tail -1 $O/k_uvic.err
python bench.py --workload uvic100_mobi21 --no-cpu-baseline > $O/k_uvic21.json 2> $O/k_uvic21.err
python bench.py --workload uvic100_ts --no-cpu-baseline > $O/k_uvic_ts.json 2> $O/k_uvic_ts.err
python bench.py > $O/k_half.json 2> $O/k_half.err; tail -1 $O/k_half.err
FP="smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"
U="python bench.py --workload uvic100_mobi37 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0"
timeout 600 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:"k_fct_march|k_update|k_invtri|k_mobi_ws|k_mobi_cell" -s 12 -c 8 -f -o $O/k_prof_uvic $U > $O/k_ncu_u.log 2>&1
python - <<'PY'
import json
for t in ("uvic", "uvic21", "uvic_ts", "half"):
    try:
        d = json.loads(open(f"gpurun_out/k_{t}.json").read().strip().splitlines()[-1])
        k = {x["kernel"]: round(x["ms_total"] / d["steps"], 3) for x in d.get("kernels", [])[:6]}
        print(t, round(d["ms_per_step"], 4), round(d["value"], 3), d["e2e"]["value"] if d.get("e2e") else None, k)
    except Exception as e:
        print(t, "failed", e)
PY
