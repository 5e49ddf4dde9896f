#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/d_pytest.log 2>&1; tail -6 $O/d_pytest.log
B="--steps 8 --no-cpu-baseline --no-e2e --min-seconds 0"
python bench.py $B > $O/d_half.json 2> $O/d_half.err
UVIC_B200_FCT_MAXW=8 python bench.py $B > $O/d_half_w8.json 2> $O/d_half_w8.err
UVIC_B200_FCT_MAXW=12 python bench.py $B > $O/d_half_w12.json 2> $O/d_half_w12.err
for v in inv5 inv6; do
  UVIC_B200_LIB=$PWD/uvic2.9_b200/variants/libuvic_b200_$v.so python bench.py $B > $O/d_half_$v.json 2> $O/d_half_$v.err
  UVIC_B200_LIB=$PWD/uvic2.9_b200/variants/libuvic_b200_$v.so python bench.py --workload uvic100_mobi37 --steps 20 --no-cpu-baseline --no-e2e > $O/d_uvic_$v.json 2> $O/d_uvic_$v.err
done
python bench.py --workload uvic100_mobi37 --steps 20 --no-cpu-baseline --no-e2e > $O/d_uvic.json 2> $O/d_uvic.err
UVIC_B200_FCT_MAXW=8 python bench.py --workload uvic100_mobi37 --steps 20 --no-cpu-baseline --no-e2e > $O/d_uvic_w8.json 2> $O/d_uvic_w8.err
UVIC_B200_FCT_MAXW=12 python bench.py --workload uvic100_mobi37 --steps 20 --no-cpu-baseline --no-e2e > $O/d_uvic_w12.json 2> $O/d_uvic_w12.err
