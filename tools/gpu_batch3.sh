#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python tools/diag_ladder.py > $O/r02_diag_ladder.log 2>&1
timeout 300 python tools/diag_ladder.py 250000 >> $O/r02_diag_ladder.log 2>&1
( timeout 2400 python -m pytest tests -m gpu -q --durations=10 > $O/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest3.log )
tail -40 $O/r02_pytest3.log | grep -v "^\[W"
cat $O/r02_diag_ladder.log
timeout 600 python tests/golden/make_golden_refcuda.py $O/golden_refcuda > $O/r02_golden_refcuda.log 2>&1; tail -3 $O/r02_golden_refcuda.log
