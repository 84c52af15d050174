#!/bin/bash
set -u
bash tools/gpu_sanitize.sh > gpurun_out/r02_sanitize_summary.log 2>&1; tail -24 gpurun_out/r02_sanitize_summary.log
