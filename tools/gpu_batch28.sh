#!/bin/bash
# body-tile rotation kernel per ladder rung; rolled free-atom ranges in part2LargeKernel
set -u
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r02_t28.log 2>&1; tail -5 $O/r02_t28.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --steps 200"
: > $O/r02_bench28.jsonl; : > $O/r02_bench28_err.log
run() { echo "# $*" >> $O/r02_bench28.jsonl; "$@" >> $O/r02_bench28.jsonl 2>> $O/r02_bench28_err.log; }
run $B --workload mixed --graph
run $B --workload mixed
run $B --workload mixed --dt-fs 2 --graph
run $B --workload mixed --dt-fs 4 --graph
run $B --workload mixed --no-fuse
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches28_mixed.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l28.log 2>&1
grep -c . $O/r02_bench28.jsonl; grep -v "^\[W" $O/r02_bench28_err.log | tail -5
