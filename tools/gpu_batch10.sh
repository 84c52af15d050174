#!/bin/bash
set -u
O=gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > $O/r02_pytest10.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest10.log )
grep -v "^\[W" $O/r02_pytest10.log | tail -6
B="timeout 400 python bench.py --steps 200 --no-cpu-baseline --no-e2e --no-gpu-reference"
: > $O/r02_bench10.jsonl; : > $O/r02_bench10_err.log
run() { echo "# $*" >> $O/r02_bench10.jsonl; "$@" >> $O/r02_bench10.jsonl 2>> $O/r02_bench10_err.log; }
run $B
RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/f/librbk.so run $B
run $B
RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/f/librbk.so run $B
run $B --layout openmm-mixed --shuffle
RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/f/librbk.so run $B --layout openmm-mixed --shuffle
run $B --dt-fs 2
run $B --dt-fs 4
grep -c . $O/r02_bench10.jsonl; grep -v "^\[W" $O/r02_bench10_err.log | tail -5
