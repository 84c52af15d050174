#!/bin/bash
set -u
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tile_stream tools/microbench/tile_stream.cu && /tmp/tile_stream > $O/r02_tile_stream.txt 2>&1
cat $O/r02_tile_stream.txt
( timeout 600 python -m pytest tests/test_glue.py -m gpu -q > $O/r02_pytest6.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest6.log )
grep -v "^\[W" $O/r02_pytest6.log | tail -8
