set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
UVIC_B200_E2E_TRACE=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/e20_uvic.json 2> $O/e20_uvic.err
grep "e2e\]" $O/e20_uvic.err | tail -3
