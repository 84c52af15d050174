#!/bin/bash
set -u
O=gpurun_out
B="timeout 400 python bench.py --workload mixed --steps 100 --no-cpu-baseline --no-e2e --no-gpu-reference"
: > $O/r02_bench12.jsonl; : > $O/r02_bench12_err.log
run() { echo "# $*" >> $O/r02_bench12.jsonl; "$@" >> $O/r02_bench12.jsonl 2>> $O/r02_bench12_err.log; }
run $B
RBK_PART2_WARP=3 run $B
RBK_PART2_WARP=4 run $B
RBK_PART2_WARP=2 RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/g/librbk.so run $B
RBK_PART2_WARP=3 RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/g/librbk.so run $B
RBK_PART2_WARP=3 timeout 600 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_parity.py -m gpu -q -k "mixed or large or huge or bitwise" > $O/r02_pytest12.log 2>&1; grep -v "^\[W" $O/r02_pytest12.log | tail -5
M="python bench.py --workload mixed --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-parity"
RBK_PART2_WARP=3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:part2WarpKernel -s 3 -c 1 -f -o $O/r02e_part2Warp_mixed $M > /dev/null 2>&1
grep -c . $O/r02_bench12.jsonl; grep -v "^\[W" $O/r02_bench12_err.log | tail -5
