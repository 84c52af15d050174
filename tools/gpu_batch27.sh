#!/bin/bash
# A/B: water headline with the library before / after the large-body work (same box, alternating)
set -u
O=gpurun_out
NEW=$PWD/openmm_rigidbody_plugin_b200/lib/librbk.so
OLD=$PWD/openmm_rigidbody_plugin_b200/lib_exp/prev_tree/openmm_rigidbody_plugin_b200/lib/librbk.so
: > $O/r02_ab27.log
for i in 1 2 3; do
  python tools/ab_step.py --lib $OLD >> $O/r02_ab27.log 2>&1
  python tools/ab_step.py --lib $NEW >> $O/r02_ab27.log 2>&1
done
python tools/ab_step.py --lib $OLD --workload mixed >> $O/r02_ab27.log 2>&1
python tools/ab_step.py --lib $NEW --workload mixed >> $O/r02_ab27.log 2>&1
grep "ms per" $O/r02_ab27.log
timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_large_bodies.py tests/test_gpu_parity.py -m gpu -q > $O/r02_t27.log 2>&1; tail -3 $O/r02_t27.log
