#!/bin/bash
set -u
O=gpurun_out
timeout 1200 python bench.py --workload mixed --graph > $O/r02_bench_config4_full.json 2> $O/r02_bench_config4_full_err.log; head -c 300 $O/r02_bench_config4_full.json; grep -v "^\[W" $O/r02_bench_config4_full_err.log | tail -3
