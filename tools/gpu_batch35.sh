#!/bin/bash
# lean water kernel: rows written by the bodies' threads (16-byte stores) - tests, A/B against the previous build
set -u
O=gpurun_out
timeout 2400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_state_changes.py -m gpu -q -x > $O/r02_t35.log 2>&1; tail -4 $O/r02_t35.log
NEW=$PWD/openmm_rigidbody_plugin_b200/lib/librbk.so
OLD=$PWD/openmm_rigidbody_plugin_b200/lib_exp/prev_tree/openmm_rigidbody_plugin_b200/lib/librbk.so
: > $O/r02_ab35.log
for i in 1 2 3; do
  python tools/ab_step.py --lib $OLD >> $O/r02_ab35.log 2>&1
  python tools/ab_step.py --lib $NEW >> $O/r02_ab35.log 2>&1
done
python tools/ab_step.py --lib $OLD --molecules 250000 >> $O/r02_ab35.log 2>&1
python tools/ab_step.py --lib $NEW --molecules 250000 >> $O/r02_ab35.log 2>&1
grep "ms per" $O/r02_ab35.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > $O/r02_b35.json 2> $O/r02_b35.err; python -c "
import json; d=json.load(open('$O/r02_b35.json')); print(d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['parity_subsample']['ok'])"
