#!/bin/bash
python -m pytest tests/test_octree_mc_gpu.py tests/test_octree_slab_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 800 compute-sanitizer --tool memcheck --print-limit 5 python scripts/dbg_slab_mc.py 2>&1 | grep -E "Invalid|at .*|by thread|ERROR SUMMARY|rank [0-9]" | head -30
