set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t8_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/t8_bench.json 2> gpurun_out/t8_bench.err
python bench.py --workload half_deg_40 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/t8_half.json 2> gpurun_out/t8_half.err
