#!/bin/bash
# round 2: k_invtri without the land stores (k_update leaves the zeros): GPU tests, 100x100x19 and 0.5 degree bench lines
set -u
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/n_pytest.log 2>&1; tail -3 $O/n_pytest.log
python bench.py --workload uvic100_mobi37 --no-cpu-baseline > $O/n_uvic.json 2> $O/n_uvic.err; tail -1 $O/n_uvic.err
python bench.py --no-cpu-baseline > $O/n_half.json 2> $O/n_half.err; tail -1 $O/n_half.err
python - <<'PY'
import json
for t in ("uvic", "half"):
    try:
        d = json.loads(open(f"gpurun_out/n_{t}.json").read().strip().splitlines()[-1])
        k = {x["kernel"]: round(x["ms_total"] / d["steps"], 3) for x in d.get("kernels", [])[:6]}
        print(t, round(d["ms_per_step"], 4), round(d["value"], 3), d["e2e"]["value"] if d.get("e2e") else None, k)
    except Exception as e:
        print(t, "failed", e)
PY
