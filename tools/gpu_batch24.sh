#!/bin/bash
# config 4: per-body columns with four interleaved partial sums (384-atom tiles, free atoms riding along)
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_openmm_layout.py tests/test_gpu_parity.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py -m gpu -q > $O/r02_t24.log 2>&1; tail -5 $O/r02_t24.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --workload mixed --steps 200"
: > $O/r02_bench24.jsonl; : > $O/r02_bench24_err.log
run() { echo "# $*" >> $O/r02_bench24.jsonl; "$@" >> $O/r02_bench24.jsonl 2>> $O/r02_bench24_err.log; }
run $B --graph
run $B
run $B --graph --layout openmm-mixed
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches24_mixed.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l24.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part2LargeKernel -s 4 -c 1 -o $O/r02i_part2Large_mixed -f python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_n24.log 2>&1; tail -2 $O/r02_n24.log
grep -c . $O/r02_bench24.jsonl; grep -v "^\[W" $O/r02_bench24_err.log | tail -5
