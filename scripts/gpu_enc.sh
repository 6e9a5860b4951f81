#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoders_gpu.py -m gpu -q --no-header -x 2>&1 | tail -12
timeout 900 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_enc.json 2> gpurun_out/bench_enc.err; tail -5 gpurun_out/bench_enc.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_enc.json').read().strip().splitlines()[-1])
print('value %.4g' % d['value'], 'mesh octree %.2f' % d['mesh_512']['octree']['latency_ms'])
print(json.dumps(d['encoders'], indent=1))
PY
