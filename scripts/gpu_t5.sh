#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_octree_mc_gpu.py tests/test_gen_mesh_flow_gpu.py tests/test_mesh_util_cpu.py -m gpu -q --no-header -x 2>&1 | tail -4
timeout 300 python scripts/mesh_latency.py 512 3 > gpurun_out/mesh_v7.json 2> gpurun_out/mesh_v7.err; tail -3 gpurun_out/mesh_v7.err; cut -c1-500 gpurun_out/mesh_v7.json
