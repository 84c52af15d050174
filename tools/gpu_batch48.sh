#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 200 --warmup 5 > $O/r02_bench_4gpu_raw.json 2> $O/r02_bench_4gpu_err.log
python tools/benchline.py < $O/r02_bench_4gpu_raw.json > $O/r02_bench_4gpu.json; head -c 400 $O/r02_bench_4gpu.json
