#!/bin/bash
# rotation kernel of large-body systems: 2 / 3 / 4 resident CTAs per SM (6 / 4 / 3 rounds of tiles on config 4)
set -u
O=gpurun_out
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --steps 200 --workload mixed --graph"
: > $O/r02_bench49.jsonl; : > $O/r02_bench49_err.log
run() { echo "# $*" >> $O/r02_bench49.jsonl; "$@" >> $O/r02_bench49.jsonl 2>> $O/r02_bench49_err.log; }
run $B
for v in rot3 rot4; do
  run env RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so $B
  run env RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so $B --dt-fs 4
  RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches49_$v.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > /dev/null 2>&1
  grep -h "part1Kernel" $O/r02_launches49_$v.csv | tail -3 | cut -c60-90,200-
done
grep -c . $O/r02_bench49.jsonl; grep -v "^\[W" $O/r02_bench49_err.log | tail -3
