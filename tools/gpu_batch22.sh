#!/bin/bash
# config 4, 384-atom tiles: free atoms riding along with an L2 prefetch a tile ahead; Part 1 as one streaming kernel
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_openmm_layout.py tests/test_gpu_parity.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py tests/test_gpu_build.py tests/test_gpu_refined.py -m gpu -q > $O/r02_t22.log 2>&1; tail -5 $O/r02_t22.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --workload mixed --steps 200"
: > $O/r02_bench22.jsonl; : > $O/r02_bench22_err.log
run() { echo "# $*" >> $O/r02_bench22.jsonl; "$@" >> $O/r02_bench22.jsonl 2>> $O/r02_bench22_err.log; }
P3=$PWD/openmm_rigidbody_plugin_b200/lib_exp/p3/librbk.so
run env RBK_LIB_PATH=$P3 $B --graph
run env RBK_LIB_PATH=$P3 $B
run env RBK_LIB_PATH=$P3 RBK_NO_STREAM_PART1=1 $B --graph
run env RBK_LIB_PATH=$P3 RBK_NO_FREE_RIDE=1 $B --graph
run env RBK_LIB_PATH=$P3 RBK_NO_FREE_RIDE=1 RBK_NO_STREAM_PART1=1 $B --graph
run env RBK_LIB_PATH=$P3 $B --graph --layout openmm-mixed
run $B --graph
RBK_LIB_PATH=$P3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches22_mixed_p3.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l22.log 2>&1
grep -c . $O/r02_bench22.jsonl; grep -v "^\[W" $O/r02_bench22_err.log | tail -5
