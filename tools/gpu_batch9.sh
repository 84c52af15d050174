#!/bin/bash
set -u
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu_err.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --impl reference --steps 10 --warmup 2 > $O/r02_bench_2gpu_ref.json 2>> $O/r02_bench_2gpu_err.log
python bench.py --steps 100 > $O/r02_bench_1gpu_b.json 2>> $O/r02_bench_2gpu_err.log
nvidia-smi topo -m > $O/r02_topo.txt 2>&1
head -c 600 $O/r02_bench_2gpu.json; echo; grep -v "^\[W\|NCCL\|^$" $O/r02_bench_2gpu_err.log | tail -5
