set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t4_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t4_bench.json 2> gpurun_out/t4_bench.err
UVIC_B200_FCT=legacy python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t4_bench_legacy.json 2> gpurun_out/t4_bench_legacy.err
python bench.py --workload half_deg_40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t4_half.json 2> gpurun_out/t4_half.err
UVIC_B200_FCT_TI=27 python bench.py --workload half_deg_40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t4_half_ti27.json 2> gpurun_out/t4_half_ti27.err
UVIC_B200_FCT_TI=16 python bench.py --workload half_deg_40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t4_half_ti16.json 2> gpurun_out/t4_half_ti16.err
