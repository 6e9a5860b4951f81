#!/bin/bash
# first GPU contact: layer kernel (simt, then tcgen05), then the query parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/$name.log; }
run layer_simt python -m pytest tests/test_query_gpu.py -q -k "layer_kernel and simt" --no-header
run layer_tc python -m pytest tests/test_query_gpu.py -q -k "layer_kernel and tcgen05" --no-header
run query_simt python -m pytest tests/test_query_gpu.py -q -k "mr_query_parity and simt" --no-header
run query_rest python -m pytest tests/test_query_gpu.py -q -k "not layer_kernel and not simt" --no-header
