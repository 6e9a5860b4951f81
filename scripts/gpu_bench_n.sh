#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value %.4g e2e %.4g ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'clocks', d['clocks'])
print('mesh', json.dumps(d.get('mesh_512')))
print('enc', json.dumps(d.get('encoders')))
PY
