#!/bin/bash
# round 2, GPU pass C (one GPU): the driver's sequence -- smoke, GPU tests, default bench line, reference arm -- plus the 100x100x19 lines
set -u
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/c_smoke.log 2>&1; tail -1 $O/c_smoke.log
python -m pytest tests -m gpu -q > $O/c_pytest.log 2>&1; tail -6 $O/c_pytest.log
python bench.py > $O/c_half.json 2> $O/c_half.err; tail -2 $O/c_half.err
python bench.py --workload uvic100_mobi37 > $O/c_uvic.json 2> $O/c_uvic.err; tail -2 $O/c_uvic.err
python bench.py --workload uvic100_mobi21 --no-cpu-baseline > $O/c_uvic21.json 2> $O/c_uvic21.err; tail -2 $O/c_uvic21.err
python bench.py --workload uvic100_ts --no-cpu-baseline > $O/c_uvic_ts.json 2> $O/c_uvic_ts.err; tail -2 $O/c_uvic_ts.err
ls $O | grep "^c_" | wc -l
