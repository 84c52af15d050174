#!/bin/bash
set -u
O=gpurun_out
./openmm_rigidbody_plugin_b200/lib_exp/tile_stream2 > $O/r02_tile_stream2.txt 2>&1; cat $O/r02_tile_stream2.txt
B="timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --steps 200"
: > $O/r02_bench29.jsonl; : > $O/r02_bench29_err.log
run() { echo "# $*" >> $O/r02_bench29.jsonl; "$@" >> $O/r02_bench29.jsonl 2>> $O/r02_bench29_err.log; }
run $B --workload mixed --graph
run $B --workload mixed --dt-fs 4 --graph
grep -c . $O/r02_bench29.jsonl; grep -v "^\[W" $O/r02_bench29_err.log | tail -5
