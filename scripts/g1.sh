set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t1_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/t1_bench_default.json 2> gpurun_out/t1_bench_default.err
python bench.py --workload half_deg_40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t1_bench_half.json 2> gpurun_out/t1_bench_half.err
ncu --set full --clock-control none --import-source on -k regex:k_gm_faces -c 1 -o gpurun_out/t1_gm_half python bench.py --workload half_deg_40 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t1_ncu.log 2>&1
