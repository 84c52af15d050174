#!/bin/bash
# free atoms riding along in part2LargeKernel: tests, config 4, launch list, ncu capture
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large_bodies.py tests/test_gpu_openmm_layout.py tests/test_gpu_parity.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py tests/test_gpu_build.py tests/test_gpu_refined.py -m gpu -q > $O/r02_t21.log 2>&1; tail -5 $O/r02_t21.log
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --workload mixed --steps 200"
: > $O/r02_bench21.jsonl; : > $O/r02_bench21_err.log
run() { echo "# $*" >> $O/r02_bench21.jsonl; "$@" >> $O/r02_bench21.jsonl 2>> $O/r02_bench21_err.log; }
P3=$PWD/openmm_rigidbody_plugin_b200/lib_exp/p3/librbk.so
run $B --graph
run $B
run env RBK_LIB_PATH=$P3 $B --graph
run env RBK_LIB_PATH=$P3 $B
run env RBK_LIB_PATH=$P3 $B --graph --layout openmm-mixed
run env RBK_LIB_PATH=$P3 $B --no-fuse
RBK_LIB_PATH=$P3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches21_mixed_p3.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l21.log 2>&1
RBK_LIB_PATH=$P3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:part2LargeKernel -s 4 -c 1 -o $O/r02g_part2Large_mixed_p3 -f python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_n21.log 2>&1; tail -2 $O/r02_n21.log
grep -c . $O/r02_bench21.jsonl; grep -v "^\[W" $O/r02_bench21_err.log | tail -5
