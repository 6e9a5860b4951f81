#!/bin/bash
# N-GPU check: sharded meshes (slab octree, dense) against the single-GPU mesh + latencies, then the bench line
N=${1:-2}
RES2=${2:-}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_gpu_check.py 512 2>&1 | grep -E "world|Error|error|Traceback" | tail -8 | tee gpurun_out/r02_dist_check_n$N.log
if [ -n "$RES2" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 scripts/dist_gpu_check.py $RES2 2>&1 | grep -E "world|Error|error|Traceback" | tail -8 | tee -a gpurun_out/r02_dist_check_n$N.log
fi
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --no-encoders > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
grep -E "Error|error|Traceback" gpurun_out/r02_bench_n$N.err | head -5
python - <<PY
import json
d = json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value %.4g e2e %.4g ms/step %.2f scaling %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['scaling']))
print('parity', json.dumps(d.get('parity')))
print('mesh', json.dumps(d.get('mesh_512')))
print('stress', json.dumps(d.get('stress_1024')))
PY
