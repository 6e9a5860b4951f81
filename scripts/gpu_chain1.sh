#!/bin/bash
# first contact of the chain kernel: smallest cases one by one under short timeouts
mkdir -p gpurun_out
for k in "2x1x128" "3x1x128" "4x8x256" "ragged" "scaled" "24x32x128"; do
  echo "=== case $k"
  timeout 120 python -m pytest tests/test_chain_gpu.py -q --no-header -x -k "test_chain_vs_oracle and $k" 2>&1 | tail -12
done
timeout 300 python -m pytest tests/test_chain_gpu.py -q --no-header -k "not test_chain_vs_oracle" 2>&1 | tail -12
