set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/e4_pytest.log
cat $O/e4_pytest.log
for W in 20 16 12; do
UVIC_B200_FCT_MAXW=$W python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/e4_uvic_w$W.json 2> $O/e4_uvic_w$W.err
UVIC_B200_FCT_MAXW=$W python bench.py --workload half_deg_40 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/e4_half_w$W.json 2> $O/e4_half_w$W.err
done
