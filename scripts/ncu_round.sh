#!/bin/bash
# Profiling pass of one round (run under gpurun): the bench lines, the launch list of the default bench command, then
# one `ncu --set full` capture of the heavy kernels on the 0.5 degree grid and on the 100x100x19 grid.
# Outputs land in gpurun_out/; scripts/ncu_traffic.py and scripts/ncu_summary.py turn them into profiles/.
set -u
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/final_uvic100.json 2> $O/final_uvic100.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/final_reference.json 2> $O/final_reference.err
python bench.py --workload half_deg_40 --steps 8 --warmup 3 > $O/final_half.json 2> $O/final_half.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > $O/plain_uvic100.json 2> $O/plain_uvic100.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_uvic100.csv $B > $O/ncu_l.log 2>&1
H="python bench.py --workload half_deg_40 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$H > $O/plain_half.json 2> $O/plain_half.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fct_march|k_update|k_invtri|k_mobi_column|k_mobi_cell" -s 10 -c 5 -f -o $O/prof_half $H > $O/ncu_h.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fct_march|k_update|k_invtri|k_mobi_ws|k_mobi_cell" -s 10 -c 5 -f -o $O/prof_uvic $B > $O/ncu_w.log 2>&1
# the momentum step on the 0.5 degree grid (k_clinic_*, k_filuv*)
python scripts/clinic_once.py > $O/clinic_once.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_clinic|k_filuv" -s 8 -c 5 -f -o $O/prof_clinic python scripts/clinic_once.py > $O/ncu_c.log 2>&1
ls -la $O | tail -20
