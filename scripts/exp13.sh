set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mobi.py -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/e13_uvic.json 2> $O/e13_uvic.err
