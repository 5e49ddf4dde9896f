set -x
O=gpurun_out
for V in 0 1; do
UVIC_B200_NVCC_EXTRA="-DFM_SKIP_F=$V" python uvic2.9_b200/build.py --force > /dev/null 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/e16_uvic_s$V.json 2> $O/e16_uvic_s$V.err
python bench.py --workload half_deg_40 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/e16_half_s$V.json 2> $O/e16_half_s$V.err
done
