#!/bin/bash
# 2 GPUs: NCCL sharded reconstruction vs single-GPU mesh, then the 2-GPU bench line
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_gpu_check.py 256 2>&1 | grep -v "^W1017\|^\*\*\*\*" | tail -8 | tee gpurun_out/dist_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 scripts/dist_gpu_check.py 512 2>&1 | grep -v "^W1017\|^\*\*\*\*" | tail -8 | tee -a gpurun_out/dist_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n$N.json
