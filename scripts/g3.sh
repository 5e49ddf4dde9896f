set -x
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t3_bench.json 2> gpurun_out/t3_bench.err
ncu --set full --clock-control none --import-source on -k regex:'k_mobi_ws|k_mobi_cell' -c 2 -o gpurun_out/t3_mobi python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t3_ncu.log 2>&1
