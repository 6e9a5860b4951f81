#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -4
timeout 300 python scripts/mesh_latency.py 512 3 > gpurun_out/mesh_v5.json 2> gpurun_out/mesh_v5.err; tail -3 gpurun_out/mesh_v5.err; cut -c1-1000 gpurun_out/mesh_v5.json
