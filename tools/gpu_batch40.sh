#!/bin/bash
set -u
O=gpurun_out
HEADLIB=$PWD/openmm_rigidbody_plugin_b200/lib_exp/head_tree/openmm_rigidbody_plugin_b200/lib/librbk.so
Q="--workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity"
RBK_LIB_PATH=$HEADLIB timeout 600 ncu --set full --clock-control none -k regex:part1Kernel -s 4 -c 1 -f -o $O/r02_rot_fixed python bench.py $Q > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:part1Kernel -s 4 -c 1 -f -o $O/r02_rot_rung0 python bench.py $Q > /dev/null 2>&1
ls -la $O/r02_rot_*.ncu-rep
