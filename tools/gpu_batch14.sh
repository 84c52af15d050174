#!/bin/bash
set -u
O=gpurun_out
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 100 --warmup 5 > $O/r02_bench_${N}gpu.json 2> $O/r02_bench_${N}gpu_err.log
nvidia-smi topo -m > $O/r02_topo_${N}gpu.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > $O/r02_lscpu.txt 2>&1
grep "^{" $O/r02_bench_${N}gpu.json | head -c 300; echo; grep -v "^\[W\|NCCL\|^$\|OMP_NUM\|\*\*\*" $O/r02_bench_${N}gpu_err.log | tail -5; cat $O/r02_lscpu.txt
