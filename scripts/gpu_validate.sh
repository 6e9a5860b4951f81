#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err; cut -c1-300 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_mesh512.csv python scripts/profile_mesh.py 512 octree > gpurun_out/prof_mesh.log 2>&1
tail -1 gpurun_out/prof_mesh.log
python scripts/launch_summary.py gpurun_out/launches_mesh512.csv 30
