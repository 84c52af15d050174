#!/bin/bash
set -u
O=gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > $O/r02_pytest11.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest11.log )
grep -v "^\[W" $O/r02_pytest11.log | tail -6
B="timeout 400 python bench.py --steps 200 --no-cpu-baseline --no-e2e --no-gpu-reference"
: > $O/r02_bench11.jsonl; : > $O/r02_bench11_err.log
run() { echo "# $*" >> $O/r02_bench11.jsonl; "$@" >> $O/r02_bench11.jsonl 2>> $O/r02_bench11_err.log; }
run $B
run $B --layout openmm-mixed --shuffle
RBK_KEEP_VELM_W=1 run $B --layout openmm-mixed --shuffle
run $B --layout openmm-double --shuffle
RBK_KEEP_VELM_W=1 run $B --layout openmm-double --shuffle
run $B --layout openmm-mixed --shuffle --dt-fs 2
N="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity"
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 6 -c 1 -o $O/r02d_fused_mode0_openmm_mixed python bench.py $N --layout openmm-mixed --shuffle > /dev/null 2>&1
grep -c . $O/r02_bench11.jsonl; grep -v "^\[W" $O/r02_bench11_err.log | tail -5
