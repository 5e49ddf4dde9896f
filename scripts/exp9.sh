set -x
O=gpurun_out
for E in 1 0; do
UVIC_B200_MOBI_EARLY=$E python bench.py --workload half_deg_40 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > $O/e9_half_early$E.json 2> $O/e9_half_early$E.err
UVIC_B200_MOBI_EARLY=$E python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > $O/e9_uvic_early$E.json 2> $O/e9_uvic_early$E.err
done
