#!/bin/bash
# atomPositionKernel: slots from per-body deltas when bodies are runs
set -u
O=gpurun_out
timeout 2400 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_openmm_layout.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py tests/test_gpu_build.py -m gpu -q -x > $O/r02_t43.log 2>&1; tail -3 $O/r02_t43.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --steps 200 --workload mixed"
: > $O/r02_bench43.jsonl; : > $O/r02_bench43_err.log
run() { echo "# $*" >> $O/r02_bench43.jsonl; "$@" >> $O/r02_bench43.jsonl 2>> $O/r02_bench43_err.log; }
run $B --graph
run $B
run $B --graph --layout openmm-mixed
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 600 ncu $M -c 40 --log-file $O/r02_launches43_mixed.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > /dev/null 2>&1
grep -c . $O/r02_bench43.jsonl; grep -v "^\[W" $O/r02_bench43_err.log | tail -3
