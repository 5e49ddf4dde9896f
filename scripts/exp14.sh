set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/e14_uvic.json 2> $O/e14_uvic.err
python bench.py --workload half_deg_40 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > $O/e14_half.json 2> $O/e14_half.err
