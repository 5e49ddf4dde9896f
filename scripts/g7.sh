set -x
python -m pytest tests/test_gpu_mobi.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t7_pytest.log
for G in 1 2 4; do
UVIC_B200_MOBI_WS_G=$G python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t7_bench_g$G.json 2> gpurun_out/t7_bench_g$G.err
done
