#!/bin/bash
python -m pytest tests/test_octree_slab_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_octree_slab_gpu.py -m gpu -x -q -k "128-16-4 and ripple" 2>&1 | grep -E "Invalid|ERROR SUMMARY|at .*\.cu|passed|failed|=========" | head -30
