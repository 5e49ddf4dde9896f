set -x
O=gpurun_out
for B in 10 12; do
UVIC_B200_NVCC_EXTRA="-DCOL_MINB=$B" python uvic2.9_b200/build.py --force > /dev/null 2>&1
python bench.py --workload half_deg_40 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/e22_half_b$B.json 2> $O/e22_half_b$B.err
done
