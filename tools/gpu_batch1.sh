#!/bin/bash
# Round-2 GPU batch 1: full GPU test-suite, bench variants, DFMA microbench, ncu of the fused kernels.
set -u
O=gpurun_out
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest1.log )
tail -5 $O/r02_pytest1.log
B="timeout 300 python bench.py --steps 100"
: > $O/r02_bench1.jsonl; : > $O/r02_bench1_err.log
run() { echo "# $*" >> $O/r02_bench1.jsonl; "$@" >> $O/r02_bench1.jsonl 2>> $O/r02_bench1_err.log; }
run $B
RBK_EAGER_FORCE_TORQUE=1 run $B --no-cpu-baseline --no-e2e
run $B --forces constant --no-cpu-baseline --no-e2e
run $B --shuffle --no-cpu-baseline --no-e2e
run $B --layout openmm-double --no-cpu-baseline --no-e2e
run $B --layout openmm-mixed --no-cpu-baseline --no-e2e
run $B --layout openmm-mixed --shuffle --no-cpu-baseline --no-e2e
run $B --dt-fs 2 --no-cpu-baseline --no-e2e
run $B --dt-fs 4 --no-cpu-baseline --no-e2e
run $B --mode 10 --no-cpu-baseline --no-e2e
run $B --workload mixed --no-cpu-baseline --no-e2e
run $B --molecules 250000 --no-cpu-baseline --no-e2e
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/dfma tools/microbench/dfma.cu && ./tools/microbench/dfma > $O/r02_dfma.txt 2>&1
N="--steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 4 -c 1 -o $O/r02_fused_mode0 python bench.py $N > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 4 -c 1 -o $O/r02_fused_mode0_openmm_mixed_shuffle python bench.py $N --layout openmm-mixed --shuffle > /dev/null 2>&1
grep -c . $O/r02_bench1.jsonl; tail -3 $O/r02_bench1_err.log
