set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fct_variants" 2>&1 | tail -30 > $O/e2_variants.log
cat $O/e2_variants.log | tail -30
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/e2_pytest.log
tail -5 $O/e2_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/e2_uvic.json 2> $O/e2_uvic.err
python bench.py --workload half_deg_40 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/e2_half.json 2> $O/e2_half.err
tail -3 $O/e2_uvic.err $O/e2_half.err
