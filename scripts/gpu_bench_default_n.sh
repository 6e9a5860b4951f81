#!/bin/bash
# the driver's own N-GPU command (no extra flags): bench.py under torchrun with the default legs, frames/s included
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_default_n$N.json 2> gpurun_out/r02_bench_default_n$N.err
echo "rc $?"
grep -E "Error|error|Traceback" gpurun_out/r02_bench_default_n$N.err | head -5
python - <<PY
import json
d = json.loads(open('gpurun_out/r02_bench_default_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value %.4g e2e %.4g ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
print('mesh', json.dumps(d.get('mesh_512')))
print('frames', json.dumps((d.get('encoders') or {}).get('frames_256_octree')))
PY
