#!/bin/bash
for ct in 148 296 592 1184 2368 4736; do
  echo -n "chunk_tiles=$ct  "
  PIFU_CHUNK_TILES=$ct timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g q/s  ms/step %.2f  gemm TF %.1f frac %.3f share %.3f avg_launch_us %.1f' % (d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step'], d['roofline']['avg_launch_us']))"
done
