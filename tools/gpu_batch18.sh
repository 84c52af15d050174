#!/bin/bash
# free atoms with two / four atoms per thread next to the register-capped large-body Part 2; tile-size variants on config 4
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_openmm_layout.py tests/test_gpu_parity.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py -m gpu -q > $O/r02_t18.log 2>&1; tail -5 $O/r02_t18.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --workload mixed --steps 200"
: > $O/r02_bench18.jsonl; : > $O/r02_bench18_err.log
run() { echo "# $*" >> $O/r02_bench18.jsonl; "$@" >> $O/r02_bench18.jsonl 2>> $O/r02_bench18_err.log; }
run $B
run $B --graph
for v in p3 p3f4 p2f4 p3r96; do
  run env RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so $B
  run env RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so $B --graph
done
run env RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/p3/librbk.so $B --layout openmm-mixed --graph
grep -c . $O/r02_bench18.jsonl; grep -v "^\[W" $O/r02_bench18_err.log | tail -5
