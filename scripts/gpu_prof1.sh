#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_octree_mc_gpu.py -q --no-header -k "device_octree" 2>&1 | tail -5
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py 151552 > gpurun_out/prof_launch.log 2>&1
tail -2 gpurun_out/prof_launch.log
# full capture of the six layer-kernel launches of the third chunk
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 12 -c 6 -o gpurun_out/gemm_r01 -f python scripts/profile_step.py 151552 > gpurun_out/prof_full.log 2>&1
tail -2 gpurun_out/prof_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 2 -c 1 -o gpurun_out/gather_r01 -f python scripts/profile_step.py 151552 > gpurun_out/prof_gather.log 2>&1
ls -la gpurun_out
