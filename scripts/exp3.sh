set -x
O=gpurun_out
H="python bench.py --workload half_deg_40 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$H > $O/e3_plain.json 2> $O/e3_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fct_march|k_invtri|k_update" -s 3 -c 3 -f -o $O/prof_march $H > $O/e3_ncu.log 2>&1
tail -3 $O/e3_ncu.log
