set -x
O=gpurun_out
H="python bench.py --workload half_deg_40 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$H > $O/e6_plain.json 2> $O/e6_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fct_march|k_update" -s 2 -c 2 -f -o $O/prof_march2 $H > $O/e6_ncu.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > $O/e6_plainu.json 2> $O/e6_plainu.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fct_march|k_update" -s 2 -c 2 -f -o $O/prof_march2u $B > $O/e6_ncuu.log 2>&1
tail -2 $O/e6_ncu.log $O/e6_ncuu.log
