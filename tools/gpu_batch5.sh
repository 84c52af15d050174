#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests/test_glue.py tests/test_gpu_constraints.py tests/test_gpu_state_changes.py tests/test_gpu_openmm_layout.py -m gpu -q > $O/r02_pytest5.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest5.log )
grep -v "^\[W" $O/r02_pytest5.log | tail -15
B="timeout 400 python bench.py --steps 200 --no-cpu-baseline --no-e2e --no-gpu-reference"
: > $O/r02_bench5.jsonl; : > $O/r02_bench5_err.log
run() { echo "# $*" >> $O/r02_bench5.jsonl; "$@" >> $O/r02_bench5.jsonl 2>> $O/r02_bench5_err.log; }
run $B
for v in a b c d e; do
  echo "# variant $v" >> $O/r02_bench5.jsonl
  RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so run $B
  RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so run $B --molecules 250000
done
grep -c . $O/r02_bench5.jsonl; grep -v "^\[W" $O/r02_bench5_err.log | tail -5
