set -x
O=gpurun_out
for G in 1 2 4; do
UVIC_B200_MOBI_WS_G=$G python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $O/e8_uvic_g$G.json 2> $O/e8_uvic_g$G.err
done
