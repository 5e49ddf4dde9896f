#!/bin/bash
# round 2, first GPU pass (one GPU): the GPU test suite, the default bench line (0.5 degree), the 100x100x19 line, the
# reference arm, then the ncu launch list and one `ncu --set full` capture of the heavy kernels of the default workload
set -u
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/a_smoke.log 2>&1; tail -1 $O/a_smoke.log
python -m pytest tests -m gpu -q -x > $O/a_pytest.log 2>&1; tail -3 $O/a_pytest.log
python bench.py > $O/a_half.json 2> $O/a_half.err; tail -2 $O/a_half.err
python bench.py --workload uvic100_mobi37 > $O/a_uvic.json 2> $O/a_uvic.err; tail -2 $O/a_uvic.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/a_ref.json 2> $O/a_ref.err; tail -2 $O/a_ref.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0"
$B > $O/a_plain_half.json 2> $O/a_plain_half.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/a_launches_half.csv $B > $O/a_ncu_l.log 2>&1
FP="smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"
timeout 1200 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:"k_fct_march|k_update|k_invtri|k_mobi_column|k_mobi_cell|k_isocoef" -s 12 -c 8 -f -o $O/a_prof_half $B > $O/a_ncu_h.log 2>&1
ls -la $O | tail -12
