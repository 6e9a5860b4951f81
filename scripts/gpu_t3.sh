#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_chain_gpu.py -m gpu -q --no-header -x 2>&1 | tail -8
timeout 300 python -m pytest tests/test_octree_mc_gpu.py tests/test_gen_mesh_flow_gpu.py -m gpu -q --no-header -x 2>&1 | tail -4
PIFU_CHAIN_TRACE=2 PIFU_CHAIN_TRACE_MIN_TILES=2000 timeout 200 python scripts/recon_phases.py 512 2>&1 | grep -E "chain trace tile [3-4]|rep 3|step 8" | tail -5
