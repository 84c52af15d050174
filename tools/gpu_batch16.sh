#!/bin/bash
# large-body Part 2 with bulk copies (part2LargeRunsKernel): tests, config 4 before/after, launch list, ncu capture
set -u
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_openmm_layout.py tests/test_gpu_parity.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py -m gpu -x -q > $O/r02_t16.log 2>&1; tail -5 $O/r02_t16.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k c4 > $O/r02_t16b.log 2>&1; tail -3 $O/r02_t16b.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --workload mixed"
: > $O/r02_bench16.jsonl; : > $O/r02_bench16_err.log
run() { echo "# $*" >> $O/r02_bench16.jsonl; "$@" >> $O/r02_bench16.jsonl 2>> $O/r02_bench16_err.log; }
run $B --steps 200
run $B --steps 200 --graph
run env RBK_NO_BULK_PART2=1 $B --steps 200
run $B --steps 200 --no-fuse
run $B --steps 200 --layout openmm-mixed --graph
run env RBK_NO_BULK_PART2=1 $B --steps 200 --layout openmm-mixed --graph
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches16_mixed.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part2LargeRunsKernel -s 4 -c 1 -o $O/r02f_part2LargeRuns_mixed -f python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_n16.log 2>&1; tail -2 $O/r02_n16.log
grep -c . $O/r02_bench16.jsonl; grep -v "^\[W" $O/r02_bench16_err.log | tail -5
