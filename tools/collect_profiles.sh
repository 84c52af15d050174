#!/bin/bash
# Run on the GPU box (gpurun -- bash tools/collect_profiles.sh): collects the round's bench lines and ncu captures
# into gpurun_out/; summarise afterwards on the CPU box with tools/ncu_summary.py and copy into profiles/.
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py > $O/${R}_bench_1gpu.json 2> $O/${R}_bench_err.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > $O/${R}_bench_reference_arm.json 2>> $O/${R}_bench_err.log
: > $O/${R}_bench_other_configs.jsonl
timeout 200 python bench.py --mode 10 --no-cpu-baseline >> $O/${R}_bench_other_configs.jsonl 2>> $O/${R}_bench_err.log
timeout 200 python bench.py --workload mixed --no-cpu-baseline >> $O/${R}_bench_other_configs.jsonl 2>> $O/${R}_bench_err.log
timeout 200 python bench.py --molecules 250000 --no-cpu-baseline >> $O/${R}_bench_other_configs.jsonl 2>> $O/${R}_bench_err.log
timeout 200 python bench.py --no-fuse --no-cpu-baseline --no-e2e >> $O/${R}_bench_other_configs.jsonl 2>> $O/${R}_bench_err.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 30 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 4 -c 1 -o $O/${R}_fused_mode0 \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Part1Kernel -s 3 -c 1 -o $O/${R}_part1_mode0 \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-fuse > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:part2Kernel -s 3 -c 1 -o $O/${R}_part2 \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-fuse > /dev/null 2>&1
# config 4 (large bodies): launch list + the two atom kernels of the large-body bucket
M="python bench.py --workload mixed --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 24 --csv --log-file $O/${R}_launches_mixed.csv $M > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:part2LargeKernel -s 3 -c 1 -f -o $O/${R}_part2Large_mixed $M > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:atomPositionKernel -s 3 -c 1 -f -o $O/${R}_atomPosition_mixed $M > /dev/null 2>&1
ls -la $O | tail -12
head -c 400 $O/${R}_bench_1gpu.json; echo
