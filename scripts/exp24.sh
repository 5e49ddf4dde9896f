set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host_buffer or lookahead or variants" 2>&1 | tail -3
UVIC_B200_E2E_TRACE=1 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $O/e24_uvic.json 2> $O/e24_uvic.err
grep "e2e\]" $O/e24_uvic.err | tail -3
