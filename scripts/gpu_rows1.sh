#!/bin/bash
# run-list chain form: targeted parity tests first, then the octree tests, the 512^3 mesh latency A/B, then the rest
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_gpu.py -m gpu -q --no-header -x 2>&1 | tail -12
timeout 600 python -m pytest tests/test_octree_mc_gpu.py tests/test_gen_mesh_flow_gpu.py -m gpu -q --no-header -x 2>&1 | tail -6
PIFU_CHAIN_ROWS=1 timeout 300 python scripts/mesh_latency.py 512 3 > gpurun_out/mesh_rows1.json 2> gpurun_out/mesh_rows1.err; tail -3 gpurun_out/mesh_rows1.err; cat gpurun_out/mesh_rows1.json
PIFU_CHAIN_ROWS=0 timeout 300 python scripts/mesh_latency.py 512 3 > gpurun_out/mesh_rows0.json 2> gpurun_out/mesh_rows0.err; tail -3 gpurun_out/mesh_rows0.err; cat gpurun_out/mesh_rows0.json
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -4
