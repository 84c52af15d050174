#!/bin/bash
# Round-2 GPU batch 2: full GPU test-suite (all failures), bench variants incl. the reference CUDA kernels, goldens, ncu.
set -u
O=gpurun_out
mkdir -p $O
( timeout 2400 python -m pytest tests -m gpu -q --durations=20 > $O/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest2.log )
tail -15 $O/r02_pytest2.log
B="timeout 400 python bench.py --steps 100"
: > $O/r02_bench2.jsonl; : > $O/r02_bench2_err.log
run() { echo "# $*" >> $O/r02_bench2.jsonl; "$@" >> $O/r02_bench2.jsonl 2>> $O/r02_bench2_err.log; }
X="--no-cpu-baseline --no-e2e --no-gpu-reference"
run $B
run $B --shuffle $X
run $B --layout openmm-mixed $X
run $B --layout openmm-mixed --shuffle $X
run $B --layout openmm-double --shuffle $X
run $B --dt-fs 2 $X
run $B --dt-fs 4 $X
run $B --forces constant $X
run $B --mode 10 --no-cpu-baseline --no-e2e
run $B --workload mixed $X
run $B --molecules 250000 $X
timeout 600 python tests/golden/make_golden_refcuda.py $O/golden_refcuda > $O/r02_golden_refcuda.log 2>&1
N="--steps 6 --warmup 3 $X --no-parity"
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 4 -c 1 -o $O/r02b_fused_mode0 python bench.py $N > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 4 -c 1 -o $O/r02b_fused_mode0_openmm_mixed python bench.py $N --layout openmm-mixed > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 4 -c 1 -o $O/r02b_fused_mode0_openmm_mixed_shuffle python bench.py $N --layout openmm-mixed --shuffle > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 30 --csv --log-file $O/r02b_launches_1M_waters_mode0.csv python bench.py --steps 10 --warmup 3 $X --no-parity > /dev/null 2>&1
grep -c . $O/r02_bench2.jsonl; tail -3 $O/r02_bench2_err.log
