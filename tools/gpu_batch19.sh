#!/bin/bash
# unified large-body Part 2 (RUNS | per-atom requests, serial sums in both): tests, config 4 with and without free atoms
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_openmm_layout.py tests/test_gpu_parity.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py tests/test_gpu_build.py tests/test_gpu_refined.py -m gpu -q > $O/r02_t19.log 2>&1; tail -5 $O/r02_t19.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --workload mixed --steps 200"
: > $O/r02_bench19.jsonl; : > $O/r02_bench19_err.log
run() { echo "# $*" >> $O/r02_bench19.jsonl; "$@" >> $O/r02_bench19.jsonl 2>> $O/r02_bench19_err.log; }
run $B --graph
run $B --graph --free-per-body 0
run env RBK_NO_BULK_PART2=1 $B --graph
run $B --graph --shuffle atoms
P3=$PWD/openmm_rigidbody_plugin_b200/lib_exp/p3/librbk.so
run env RBK_LIB_PATH=$P3 $B --graph
run env RBK_LIB_PATH=$P3 $B --graph --free-per-body 0
run env RBK_LIB_PATH=$P3 $B --graph --layout openmm-mixed
run env RBK_LIB_PATH=$P3 RBK_NO_BULK_PART2=1 $B --graph
grep -c . $O/r02_bench19.jsonl; grep -v "^\[W" $O/r02_bench19_err.log | tail -5
