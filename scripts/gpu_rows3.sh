#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py tests/test_octree_mc_gpu.py tests/test_gen_mesh_flow_gpu.py tests/test_mesh_util_cpu.py -m gpu -q --no-header -x 2>&1 | tail -6
timeout 300 python scripts/mesh_latency.py 512 3 > gpurun_out/mesh_v3.json 2> gpurun_out/mesh_v3.err; tail -3 gpurun_out/mesh_v3.err; cat gpurun_out/mesh_v3.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_mesh512.csv python scripts/profile_mesh.py 512 octree > gpurun_out/prof_mesh.log 2>&1
tail -2 gpurun_out/prof_mesh.log
python scripts/launch_summary.py gpurun_out/launches_mesh512.csv 24
