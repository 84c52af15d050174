#!/bin/bash
# ladder epilogue without the unconditional fence: water A/B, rotation kernel time, tests
set -u
O=gpurun_out
NEW=$PWD/openmm_rigidbody_plugin_b200/lib/librbk.so
OLD=$PWD/openmm_rigidbody_plugin_b200/lib_exp/prev_tree/openmm_rigidbody_plugin_b200/lib/librbk.so
: > $O/r02_ab39.log
for i in 1 2 3; do
  python tools/ab_step.py --lib $OLD >> $O/r02_ab39.log 2>&1
  python tools/ab_step.py --lib $NEW >> $O/r02_ab39.log 2>&1
done
python tools/ab_step.py --lib $OLD --molecules 250000 >> $O/r02_ab39.log 2>&1
python tools/ab_step.py --lib $NEW --molecules 250000 >> $O/r02_ab39.log 2>&1
grep "ms per" $O/r02_ab39.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches39_mixed.csv python bench.py --workload mixed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_l39.log 2>&1
grep -h "part1Kernel" $O/r02_launches39_mixed.csv | tail -4 | cut -c1-60,200-
timeout 2400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_large_bodies.py -m gpu -q -x > $O/r02_t39.log 2>&1; tail -3 $O/r02_t39.log
