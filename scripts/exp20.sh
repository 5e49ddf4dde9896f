set -x
O=gpurun_out
UVIC_B200_E2E_TRACE=1 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > $O/e20_uvic.json 2> $O/e20_uvic.err
grep "e2e\]" $O/e20_uvic.err | tail -8
