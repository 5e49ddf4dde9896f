set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/e5_pytest.log
cat $O/e5_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/e5_uvic.json 2> $O/e5_uvic.err
python bench.py --workload half_deg_40 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/e5_half.json 2> $O/e5_half.err
