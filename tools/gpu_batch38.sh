#!/bin/bash
set -u
O=gpurun_out
for v in r0plain; do
RBK_LIB_PATH=$PWD/openmm_rigidbody_plugin_b200/lib_exp/$v/librbk.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches38_$v.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l38.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches38_cur.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l38.log 2>&1
grep -h "part1Kernel" $O/r02_launches38_r0plain.csv | tail -3 | cut -c1-60,200-
grep -h "part1Kernel" $O/r02_launches38_cur.csv | tail -3 | cut -c1-60,200-
