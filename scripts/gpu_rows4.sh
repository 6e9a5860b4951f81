#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py tests/test_octree_mc_gpu.py tests/test_gen_mesh_flow_gpu.py -m gpu -q --no-header -x 2>&1 | tail -4
timeout 300 python scripts/mesh_latency.py 512 3 > gpurun_out/mesh_v4.json 2> gpurun_out/mesh_v4.err; tail -3 gpurun_out/mesh_v4.err; cut -c1-900 gpurun_out/mesh_v4.json
timeout 300 python scripts/recon_phases.py 512 2>&1 | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chain_kernel" --launch-skip 16 -c 1 -o gpurun_out/chain_rows_r01b -f python scripts/profile_mesh.py 512 octree > gpurun_out/prof_rows_full.log 2>&1
tail -2 gpurun_out/prof_rows_full.log
