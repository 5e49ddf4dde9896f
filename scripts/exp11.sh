set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/e11_pytest.log
cat $O/e11_pytest.log | tail -15
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/e11_uvic.json 2> $O/e11_uvic.err
tail -5 $O/e11_uvic.err
