#!/bin/bash
# compute-sanitizer over the small-system GPU tests that exercise every kernel variant (memcheck + racecheck + synccheck)
set -u
O=gpurun_out
T="tests/test_gpu_fused.py tests/test_gpu_openmm_layout.py tests/test_gpu_large_bodies.py tests/test_gpu_state_changes.py tests/test_gpu_constraints.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest $T -m gpu -q -x > $O/r02_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r02_sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 77 python -m pytest tests/test_gpu_fused.py tests/test_gpu_openmm_layout.py tests/test_gpu_large_bodies.py -m gpu -q -x > $O/r02_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r02_sanitize_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 77 python -m pytest tests/test_gpu_fused.py tests/test_gpu_large_bodies.py -m gpu -q -x > $O/r02_sanitize_synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/r02_sanitize_synccheck.log
for f in memcheck racecheck synccheck; do echo "== $f"; grep -v "^\[W" $O/r02_sanitize_$f.log | tail -8; done
