#!/bin/bash
set -u
O=gpurun_out
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference"
: > $O/r02_bench15.jsonl; : > $O/r02_bench15_err.log
run() { echo "# $*" >> $O/r02_bench15.jsonl; "$@" >> $O/r02_bench15.jsonl 2>> $O/r02_bench15_err.log; }
run $B --steps 3000
run $B --steps 2000 --dt-fs 2
run $B --steps 400 --layout openmm-mixed --shuffle --graph
run $B --steps 400 --workload mixed --layout openmm-mixed
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; tail -3 $O/r02_smoke.log
grep -c . $O/r02_bench15.jsonl; grep -v "^\[W" $O/r02_bench15_err.log | tail -5
