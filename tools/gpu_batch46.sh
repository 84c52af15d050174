#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_constraints.py -m gpu -q > gpurun_out/r02_t46.log 2>&1; tail -4 gpurun_out/r02_t46.log
