#!/bin/bash
# state check of the committed v2 layer kernel: parity tests, bench for both CTA modes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -15
for impl in pair tc1; do
  echo -n "impl=$impl  "
  PIFU_GEMM_IMPL=$impl timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$impl.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g q/s  ms/step %.2f  e2e %.4g gemm TF %.1f frac %.3f share %.3f avg_launch_us %.1f clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step'], d['roofline']['avg_launch_us'], d['clocks']))"
done
