set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $O/e24_uvic.json 2> $O/e24_uvic.err
python bench.py --workload half_deg_40 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > $O/e24_half.json 2> $O/e24_half.err
