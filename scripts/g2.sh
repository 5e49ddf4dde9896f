set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t2_pytest.log
for ws in 0 1; do
UVIC_B200_MOBI_WS=$ws python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t2_bench_ws$ws.json 2> gpurun_out/t2_bench_ws$ws.err
done
UVIC_B200_MOBI_WS=0 python bench.py --workload half_deg_40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t2_half_ws0.json 2> gpurun_out/t2_half_ws0.err
UVIC_B200_MOBI_WS=1 python bench.py --workload half_deg_40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t2_half_ws1.json 2> gpurun_out/t2_half_ws1.err
