#!/bin/bash
set -u
O=gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > $O/r02_pytest13.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest13.log )
grep -v "^\[W" $O/r02_pytest13.log | tail -6
B="timeout 400 python bench.py --steps 200 --no-cpu-baseline --no-e2e --no-gpu-reference"
: > $O/r02_bench13.jsonl; : > $O/r02_bench13_err.log
run() { echo "# $*" >> $O/r02_bench13.jsonl; "$@" >> $O/r02_bench13.jsonl 2>> $O/r02_bench13_err.log; }
run $B
run $B --layout soa
run $B --layout soa --shuffle
run $B --layout openmm-mixed --shuffle
run $B --mode 10
grep -c . $O/r02_bench13.jsonl; grep -v "^\[W" $O/r02_bench13_err.log | tail -5
