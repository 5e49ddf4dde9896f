set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for B in 1 3 4; do
UVIC_B200_NVCC_EXTRA="-DUPD_MINB=$B" python uvic2.9_b200/build.py --force > /dev/null 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/e10_uvic_b$B.json 2> $O/e10_uvic_b$B.err
python bench.py --workload half_deg_40 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/e10_half_b$B.json 2> $O/e10_half_b$B.err
done
