#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -6
bash scripts/gpu_mesh1.sh
