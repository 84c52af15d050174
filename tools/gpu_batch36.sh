#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu_err.log
tail -c 600 $O/r02_bench_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 2 > $O/r02_bench_2gpu_reference_arm.json 2>> $O/r02_bench_2gpu_err.log
tail -c 300 $O/r02_bench_2gpu_reference_arm.json
