#!/bin/bash
python -m pytest tests/test_octree_mc_gpu.py -m gpu -x -q 2>&1 | tail -3
for wl in 12 8; do echo "WL=$wl"; PIFU_MC_WL=$wl python scripts/profile_mc.py 512 5; done
PIFU_MC_TEAM=64 python scripts/profile_mc.py 512 5
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_mc_v8_launches.csv python scripts/profile_mc.py 512 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_mc_v8_launches.csv 40 | grep -E "total|classify|emit|scan|active"
