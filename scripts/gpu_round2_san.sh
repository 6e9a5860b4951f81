#!/bin/bash
# compute-sanitizer memcheck over the kernels that are new in round 2 (marching cubes, refine, meshclean, slab octree, split layer kernel)
run() { echo "== $*"; timeout 900 compute-sanitizer --tool memcheck --print-limit 3 "$@" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|at pifu" | head -8; }
run python -m pytest tests/test_octree_mc_gpu.py -m gpu -x -q -k "marching or reconstruction_end_to_end"
run python -m pytest tests/test_postprocess_gpu.py tests/test_octree_slab_gpu.py -m gpu -x -q -k "not 256-32-8"
run python -m pytest tests/test_precision_gpu.py -m gpu -x -q -k "hybrid_points or hybrid_chain or k_concatenated"
run python -m pytest tests/test_coarse_only_gpu.py -m gpu -x -q -k "ragged"
