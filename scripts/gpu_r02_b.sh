#!/bin/bash
# round 2, second GPU pass: GPU tests with the TMA-staged march, bench lines, A/B of FMA contraction and of the TMA staging
set -u
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/b_smoke.log 2>&1; tail -1 $O/b_smoke.log
python -m pytest tests -m gpu -q > $O/b_pytest.log 2>&1; tail -8 $O/b_pytest.log
B="--steps 8 --no-cpu-baseline --no-e2e --min-seconds 0"
python bench.py $B > $O/b_half.json 2> $O/b_half.err; tail -2 $O/b_half.err
UVIC_B200_FCT_TMA=0 python bench.py $B > $O/b_half_notma.json 2> $O/b_half_notma.err
python bench.py --workload uvic100_mobi37 --steps 20 --no-cpu-baseline --no-e2e > $O/b_uvic.json 2> $O/b_uvic.err; tail -2 $O/b_uvic.err
UVIC_B200_FCT_TMA=0 python bench.py --workload uvic100_mobi37 --steps 20 --no-cpu-baseline --no-e2e > $O/b_uvic_notma.json 2> $O/b_uvic_notma.err
for v in fmad_fct fmad_mobi fmad_tracer fmad_all; do
  if [ -f uvic2.9_b200/variants/libuvic_b200_$v.so ]; then
    UVIC_B200_LIB=$PWD/uvic2.9_b200/variants/libuvic_b200_$v.so python bench.py $B > $O/b_half_$v.json 2> $O/b_half_$v.err
    UVIC_B200_LIB=$PWD/uvic2.9_b200/variants/libuvic_b200_$v.so python bench.py --workload uvic100_mobi37 --steps 20 --no-cpu-baseline --no-e2e > $O/b_uvic_$v.json 2> $O/b_uvic_$v.err
  fi
done
UVIC_B200_LIB=$PWD/uvic2.9_b200/variants/libuvic_b200_fmad_all.so python -m pytest tests -m gpu -q > $O/b_pytest_fmad_all.log 2>&1; tail -8 $O/b_pytest_fmad_all.log
ls $O | grep "^b_" | wc -l
