#!/bin/bash
# part2LargeRunsKernel with compile-time stage capacity: tests, then tile-size / residency variants on config 4
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_openmm_layout.py tests/test_gpu_parity.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py tests/test_gpu_fused.py -m gpu -q > $O/r02_t17.log 2>&1; tail -5 $O/r02_t17.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --workload mixed --steps 200"
: > $O/r02_bench17.jsonl; : > $O/r02_bench17_err.log
run() { echo "# $*" >> $O/r02_bench17.jsonl; "$@" >> $O/r02_bench17.jsonl 2>> $O/r02_bench17_err.log; }
run $B
run $B --graph
for v in m6 m4 p3 p4; do
  run env RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so $B
  run env RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so $B --graph
done
run $B --layout openmm-mixed --graph
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches17_mixed.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l17.log 2>&1
for v in p3 p4; do
RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches17_mixed_$v.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l17.log 2>&1
done
grep -c . $O/r02_bench17.jsonl; grep -v "^\[W" $O/r02_bench17_err.log | tail -5
