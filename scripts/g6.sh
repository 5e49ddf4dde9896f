set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t6_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t6_bench.json 2> gpurun_out/t6_bench.err
UVIC_B200_FCT=merged python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t6_bench_merged.json 2> gpurun_out/t6_bench_merged.err
python bench.py --workload half_deg_40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t6_half.json 2> gpurun_out/t6_half.err
UVIC_B200_FCT=merged python bench.py --workload half_deg_40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t6_half_merged.json 2> gpurun_out/t6_half_merged.err
