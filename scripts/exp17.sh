set -x
O=gpurun_out
for G in 1 4; do
UVIC_B200_MOBI_WS=1 UVIC_B200_MOBI_WS_G=$G python bench.py --workload half_deg_40 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/e17_half_ws$G.json 2> $O/e17_half_ws$G.err
done
UVIC_B200_MOBI_WS=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/e17_uvic_col.json 2> $O/e17_uvic_col.err
