#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 2400 python -m pytest tests -m gpu -q --durations=10 > $O/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest4.log )
grep -v "^\[W" $O/r02_pytest4.log | tail -25
B="timeout 400 python bench.py --steps 100"
: > $O/r02_bench4.jsonl; : > $O/r02_bench4_err.log
run() { echo "# $*" >> $O/r02_bench4.jsonl; "$@" >> $O/r02_bench4.jsonl 2>> $O/r02_bench4_err.log; }
X="--no-cpu-baseline --no-e2e --no-gpu-reference"
run $B
run $B --shuffle $X
run $B --layout openmm-mixed $X
run $B --layout openmm-mixed --shuffle $X
run $B --layout openmm-double --shuffle $X
run $B --mode 10 --no-cpu-baseline --no-e2e
run $B --molecules 250000 $X
run $B --no-fuse $X
grep -c . $O/r02_bench4.jsonl; grep -v "^\[W" $O/r02_bench4_err.log | tail -5
