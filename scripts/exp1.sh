set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/e1_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/e1_uvic.json 2> $O/e1_uvic.err
python bench.py --workload half_deg_40 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/e1_half.json 2> $O/e1_half.err
for G in 1 2 4; do
UVIC_B200_MOBI_WS=1 UVIC_B200_MOBI_WS_G=$G python bench.py --workload half_deg_40 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/e1_half_ws$G.json 2> $O/e1_half_ws$G.err
done
UVIC_B200_FMAD=k_tracer.cu python uvic2.9_b200/build.py --force
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/e1_pytest_fmad.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/e1_uvic_fmad.json 2> $O/e1_uvic_fmad.err
python bench.py --workload half_deg_40 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/e1_half_fmad.json 2> $O/e1_half_fmad.err
cat $O/e1_pytest.log $O/e1_pytest_fmad.log
