#!/bin/bash
# Round-2 evidence, collected on the GPU box in one call (gpurun -- bash tools/collect_profiles_r02.sh): every committed bench
# line with the command that produced it, launch lists with DRAM bytes per kernel, the ncu captures of the kernels that changed
# last.  Summarise on the CPU box with tools/ncu_summary.py and copy into profiles/.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu_err.log
B="timeout 600 python bench.py"
Q="--no-cpu-baseline --no-e2e --no-gpu-reference"
: > $O/r02_bench_all_configs.jsonl; : > $O/r02_bench_all_err.log
run() { echo "# $*" >> $O/r02_bench_all_configs.jsonl; "$@" >> $O/r02_bench_all_configs.jsonl 2>> $O/r02_bench_all_err.log; }
run $B --impl reference --steps 20 --warmup 2
run $B --forces constant --steps 100 $Q
run $B --graph $Q
run $B --mode 10 --no-cpu-baseline --no-e2e
run $B --mode 10 --graph $Q
run $B --workload mixed $Q
run $B --workload mixed --graph $Q
run $B --workload mixed --no-fuse $Q
run $B --workload mixed --graph --dt-fs 2 $Q
run $B --workload mixed --graph --dt-fs 4 $Q
run $B --workload mixed --graph --layout openmm-mixed $Q
run $B --workload mixed --graph --shuffle $Q
run $B --workload mixed --graph --free-per-body 0 $Q
run $B --molecules 250000 $Q
run $B --molecules 250000 --graph $Q
run $B --dt-fs 2 $Q
run $B --dt-fs 4 $Q
run $B --layout soa $Q
run $B --shuffle $Q
run $B --layout openmm-mixed --shuffle $Q
run $B --layout openmm-double --shuffle $Q
run $B --layout openmm-mixed --shuffle atoms $Q
run $B --no-fuse $Q
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 600 ncu $M -c 40 --log-file $O/r02_launches_mixed_config4.csv python bench.py --workload mixed --steps 6 --warmup 3 $Q --no-parity > /dev/null 2>&1
timeout 600 ncu $M -c 30 --log-file $O/r02_launches_1M_waters_mode0.csv python bench.py --steps 10 --warmup 3 $Q --no-parity > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part2LargeKernel -s 4 -c 1 -f -o $O/r02_part2Large_mixed python bench.py --workload mixed --steps 6 --warmup 3 $Q --no-parity > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:atomPositionKernel -s 4 -c 1 -f -o $O/r02_atomPosition_mixed python bench.py --workload mixed --steps 6 --warmup 3 $Q --no-parity > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; tail -2 $O/r02_smoke.log
grep -c . $O/r02_bench_all_configs.jsonl; grep -v "^\[W" $O/r02_bench_all_err.log | tail -5; head -c 300 $O/r02_bench_1gpu.json
