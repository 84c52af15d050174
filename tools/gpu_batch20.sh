#!/bin/bash
# free atoms: side stream next to the body kernels vs a plain launch in front of them
set -u
O=gpurun_out
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --workload mixed --steps 200"
: > $O/r02_bench20.jsonl; : > $O/r02_bench20_err.log
run() { echo "# $*" >> $O/r02_bench20.jsonl; "$@" >> $O/r02_bench20.jsonl 2>> $O/r02_bench20_err.log; }
P3=$PWD/openmm_rigidbody_plugin_b200/lib_exp/p3/librbk.so
run env RBK_NO_SIDE_STREAM=1 $B --graph
run env RBK_NO_SIDE_STREAM=1 $B
run env RBK_NO_SIDE_STREAM=1 RBK_LIB_PATH=$P3 $B --graph
run env RBK_NO_SIDE_STREAM=1 RBK_LIB_PATH=$P3 $B
run env RBK_NO_SIDE_STREAM=1 RBK_LIB_PATH=$P3 $B --graph --layout openmm-mixed
run env RBK_NO_SIDE_STREAM=1 RBK_LIB_PATH=$P3 $B --no-fuse
grep -c . $O/r02_bench20.jsonl; grep -v "^\[W" $O/r02_bench20_err.log | tail -5
