#!/bin/bash
# measurement batch: every line that goes into profiles/r02_*.json(l)
set -u
O=gpurun_out
B="timeout 600 python bench.py"
: > $O/r02_bench8.jsonl; : > $O/r02_bench8_err.log
run() { echo "# $*" >> $O/r02_bench8.jsonl; "$@" >> $O/r02_bench8.jsonl 2>> $O/r02_bench8_err.log; }
X="--no-cpu-baseline --no-e2e --no-gpu-reference"
run $B
run $B --impl reference --steps 20 --warmup 2
run $B --forces constant --steps 100 $X
run $B --graph $X
run $B --mode 10 --no-cpu-baseline --no-e2e
run $B --mode 10 --graph $X
run $B --workload mixed $X
run $B --workload mixed --graph $X
run $B --molecules 250000 $X
run $B --molecules 250000 --graph $X
run $B --dt-fs 2 $X
run $B --dt-fs 4 $X
run $B --layout soa $X
run $B --shuffle $X
run $B --layout openmm-mixed --shuffle $X
run $B --layout openmm-double --shuffle $X
run $B --layout openmm-mixed --shuffle atoms $X
run $B --no-fuse $X
N="--steps 6 --warmup 3 $X --no-parity"
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 6 -c 1 -o $O/r02c_fused_mode0 python bench.py $N > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 6 -c 1 -o $O/r02c_fused_mode0_openmm_mixed python bench.py $N --layout openmm-mixed --shuffle > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 6 -c 1 -o $O/r02c_fused_mode10 python bench.py $N --mode 10 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 40 --csv --log-file $O/r02c_launches_1M_waters_mode0.csv python bench.py --steps 10 --warmup 3 $X --no-parity > /dev/null 2>&1
M="python bench.py --workload mixed --steps 4 --warmup 3 $X --no-parity"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 30 --csv --log-file $O/r02c_launches_mixed.csv $M > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:part2LargeKernel -s 3 -c 1 -f -o $O/r02c_part2Large_mixed $M > /dev/null 2>&1
grep -c . $O/r02_bench8.jsonl; grep -v "^\[W" $O/r02_bench8_err.log | tail -5
