#!/bin/bash
# chain kernel v3: full parity suite, bench (full line), ncu launch list + full capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err; cut -c1-400 gpurun_out/bench_full.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chain.csv python scripts/profile_step.py 2424832 > gpurun_out/prof_launch.log 2>&1
tail -1 gpurun_out/prof_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -o gpurun_out/chain_r01_v3 -f python scripts/profile_step.py 2424832 > gpurun_out/prof_full.log 2>&1
tail -2 gpurun_out/prof_full.log
