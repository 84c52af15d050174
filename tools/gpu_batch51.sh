#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_large_bodies.py -m gpu -q > gpurun_out/r02_t51.log 2>&1; tail -4 gpurun_out/r02_t51.log
