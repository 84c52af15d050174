#!/bin/bash
set -u
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_t50.log 2>&1; tail -3 $O/r02_t50.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final_err.log; head -c 500 $O/r02_bench_final.json
