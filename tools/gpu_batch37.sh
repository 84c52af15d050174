#!/bin/bash
# body-tile rotation kernel on a 12 / 13 / 16 ladder: tests, config 4 at 1 / 2 / 4 / 5 fs, launch list
set -u
O=gpurun_out
timeout 2400 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_state_changes.py tests/test_gpu_openmm_layout.py tests/test_gpu_fused.py -m gpu -q -x > $O/r02_t37.log 2>&1; tail -4 $O/r02_t37.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --steps 200 --workload mixed --graph"
: > $O/r02_bench37.jsonl; : > $O/r02_bench37_err.log
run() { echo "# $*" >> $O/r02_bench37.jsonl; "$@" >> $O/r02_bench37.jsonl 2>> $O/r02_bench37_err.log; }
run $B
run $B --dt-fs 2
run $B --dt-fs 4
run $B --dt-fs 5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches37_mixed.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l37.log 2>&1
grep -c . $O/r02_bench37.jsonl; grep -v "^\[W" $O/r02_bench37_err.log | tail -5
