#!/bin/bash
# round 2: strong scaling of the 0.5 degree workload with the final kernels, N = $1 (torchrun), bitwise N-slab check in the run
set -u
N=$1; O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu-baseline > $O/l_half_n$N.json 2> $O/l_half_n$N.err
tail -2 $O/l_half_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/l_half_n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("multi_gpu_parity"))
PY
