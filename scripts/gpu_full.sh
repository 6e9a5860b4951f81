#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q --no-header -m gpu 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_full.json; python -c "
import json
d=json.load(open('gpurun_out/bench_full.json'))
print('value %.4g q/s  ms/step %.2f e2e %.4g gemm TF %.1f frac %.3f cpu %.4g cores %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']))"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
