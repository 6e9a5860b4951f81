#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run query python -m pytest tests/test_query_gpu.py -q --no-header -x
TAILN=40 run octmc python -m pytest tests/test_octree_mc_gpu.py -q --no-header
run smoke python __graft_entry__.py smoke
TAILN=5 run bench python bench.py --steps 3 --warmup 3
