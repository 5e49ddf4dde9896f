#!/bin/bash
# what the driver runs at round end, in one go (one GPU): smoke, the GPU test suite, the default bench line
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > $O/check_default.json 2> $O/check_default.err; tail -1 $O/check_default.err
