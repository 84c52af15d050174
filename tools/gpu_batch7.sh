#!/bin/bash
set -u
O=gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q --durations=5 > $O/r02_pytest7.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest7.log )
grep -v "^\[W" $O/r02_pytest7.log | tail -12
B="timeout 400 python bench.py --steps 200 --no-cpu-baseline --no-e2e --no-gpu-reference"
: > $O/r02_bench7.jsonl; : > $O/r02_bench7_err.log
run() { echo "# $*" >> $O/r02_bench7.jsonl; "$@" >> $O/r02_bench7.jsonl 2>> $O/r02_bench7_err.log; }
run $B
RBK_FULL_LADDER=1 run $B
run $B --molecules 250000
run $B --dt-fs 2
run $B --dt-fs 4
run $B --layout openmm-mixed --shuffle
run $B --forces constant
grep -c . $O/r02_bench7.jsonl; grep -v "^\[W" $O/r02_bench7_err.log | tail -5
