#!/bin/bash
set -u
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_t31.log 2>&1; tail -3 $O/r02_t31.log
bash tools/collect_profiles_r02.sh > $O/r02_collect31.log 2>&1; tail -4 $O/r02_collect31.log
