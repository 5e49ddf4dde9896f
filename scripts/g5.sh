set -x
python -m pytest tests/test_gpu_mobi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t5_pytest.log
UVIC_B200_FCT=legacy python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t5_bench.json 2> gpurun_out/t5_bench.err
UVIC_B200_FCT=legacy ncu --set full --clock-control none --import-source on -k regex:'k_mobi_ws' -s 2 -c 1 -o gpurun_out/t5_mobi python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t5_ncu.log 2>&1
