#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_octree_mc_gpu.py tests/test_gen_mesh_flow_gpu.py tests/test_mesh_util_cpu.py tests/test_chain_gpu.py -m gpu -q --no-header -x 2>&1 | tail -4
timeout 300 python scripts/mesh_latency.py 512 3 > gpurun_out/mesh_v6.json 2> gpurun_out/mesh_v6.err; tail -3 gpurun_out/mesh_v6.err; cut -c1-1100 gpurun_out/mesh_v6.json
timeout 300 python scripts/recon_phases.py 512 2>&1 | tail -4
