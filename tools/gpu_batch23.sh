#!/bin/bash
# config 4: chunked sums in phase B2 (384-atom tiles, free atoms riding along)
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_openmm_layout.py tests/test_gpu_parity.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py tests/test_gpu_build.py tests/test_gpu_refined.py -m gpu -q > $O/r02_t23.log 2>&1; tail -5 $O/r02_t23.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --workload mixed --steps 200"
: > $O/r02_bench23.jsonl; : > $O/r02_bench23_err.log
run() { echo "# $*" >> $O/r02_bench23.jsonl; "$@" >> $O/r02_bench23.jsonl 2>> $O/r02_bench23_err.log; }
run $B --graph
run $B
run env RBK_NO_FREE_RIDE=1 $B --graph
run $B --graph --layout openmm-mixed
run $B --graph --free-per-body 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches23_mixed.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l23.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part2LargeKernel -s 4 -c 1 -o $O/r02h_part2Large_mixed -f python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_n23.log 2>&1; tail -2 $O/r02_n23.log
grep -c . $O/r02_bench23.jsonl; grep -v "^\[W" $O/r02_bench23_err.log | tail -5
