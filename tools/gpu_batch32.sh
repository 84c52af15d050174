#!/bin/bash
set -u
O=gpurun_out
BENCH_DEBUG_TF=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > $O/r02_tf32.json 2> $O/r02_tf32.log
grep "\[bench\]" $O/r02_tf32.log
python -c "
import json; d=json.load(open('$O/r02_tf32.json')); print(d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'])"
python tools/ab_step.py --lib $PWD/openmm_rigidbody_plugin_b200/lib/librbk.so
BENCH_DEBUG_TF=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity --steps 1000 > $O/r02_tf32b.json 2> $O/r02_tf32b.log
grep "\[bench\]" $O/r02_tf32b.log
python -c "
import json; d=json.load(open('$O/r02_tf32b.json')); print(d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'])"
