#!/bin/bash
# soak: long runs re-checked against the oracle on a subsample (median gate beyond 1200 steps)
set -u
O=gpurun_out
B="timeout 1500 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference"
: > $O/r02_soak47.jsonl; : > $O/r02_soak47_err.log
run() { echo "# $*" >> $O/r02_soak47.jsonl; "$@" >> $O/r02_soak47.jsonl 2>> $O/r02_soak47_err.log; }
run $B --steps 3000
run $B --workload mixed --steps 1500
run $B --workload mixed --steps 1000 --layout openmm-mixed --graph
run $B --workload mixed --steps 300 --dt-fs 2
grep -c . $O/r02_soak47.jsonl; grep -v "^\[W" $O/r02_soak47_err.log | tail -4 | cut -c1-300
