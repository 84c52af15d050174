#!/bin/bash
# full GPU suite + compute-sanitizer on the new large-body kernel + bench lines for profiles/
set -u
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_t25.log 2>&1; tail -5 $O/r02_t25.log
bash tools/gpu_sanitize.sh > $O/r02_sanitize_summary.log 2>&1; tail -30 $O/r02_sanitize_summary.log
