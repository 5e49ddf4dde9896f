#!/bin/bash
# round 2, final single-GPU pass: what the driver runs (smoke, GPU tests, default bench line, reference arm) plus the
# profiling evidence of the final kernels: ncu launch list of the default bench command and one `ncu --set full` capture of its
# heavy kernels (+ FP64 instruction counts) on the 0.5 degree and on the 100x100x19 workload
set -u
O=gpurun_out; P=${PASS:-e}
python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1; tail -1 $O/${P}_smoke.log
python -m pytest tests -m gpu -q > $O/${P}_pytest.log 2>&1; tail -4 $O/${P}_pytest.log
python bench.py > $O/${P}_half.json 2> $O/${P}_half.err; tail -2 $O/${P}_half.err
python bench.py --impl reference > $O/${P}_ref.json 2> $O/${P}_ref.err; tail -2 $O/${P}_ref.err
python bench.py --workload uvic100_mobi37 > $O/${P}_uvic.json 2> $O/${P}_uvic.err; tail -2 $O/${P}_uvic.err
python bench.py --workload uvic100_mobi21 --no-cpu-baseline > $O/${P}_uvic21.json 2> $O/${P}_uvic21.err
python bench.py --workload uvic100_ts --no-cpu-baseline > $O/${P}_uvic_ts.json 2> $O/${P}_uvic_ts.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0"
$B > $O/${P}_plain_half.json 2> $O/${P}_plain_half.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${P}_launches_half.csv $B > $O/${P}_ncu_l.log 2>&1
FP="smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"
timeout 1200 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:"k_fct_march|k_update|k_invtri|k_mobi_column|k_mobi_cell|k_isocoef" -s 12 -c 8 -f -o $O/${P}_prof_half $B > $O/${P}_ncu_h.log 2>&1
U="python bench.py --workload uvic100_mobi37 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0"
timeout 600 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:"k_fct_march|k_update|k_invtri|k_mobi_ws|k_mobi_cell" -s 12 -c 8 -f -o $O/${P}_prof_uvic $U > $O/${P}_ncu_u.log 2>&1
UVIC_B200_FCT_CHUNK=360 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0.3 > $O/${P}_chunk360.json 2> $O/${P}_chunk360.err
ls -la $O | grep " ${P}_" | wc -l
