#!/bin/bash
set -u
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_t45.log 2>&1; tail -4 gpurun_out/r02_t45.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
