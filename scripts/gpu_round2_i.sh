#!/bin/bash
N=${1:-8}
CUDA_LAUNCH_BLOCKING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_gpu_check.py 512 > gpurun_out/r02_dbg_n$N.log 2>&1
grep -E "world|rank[0-9]\]:   File|rank[0-9]\]:     |Error|error:|PifuError" gpurun_out/r02_dbg_n$N.log | head -60 | cut -c1-250
