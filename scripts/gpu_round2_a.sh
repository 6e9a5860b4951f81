#!/bin/bash
# GPU validation of the marching-cubes rewrite + bench line (round 2)
python -m pytest tests/test_octree_mc_gpu.py tests/test_fullsize_gpu.py tests/test_gen_mesh_flow_gpu.py tests/test_coarse_only_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_mc_tests.log
cat gpurun_out/r02_mc_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err
tail -c 2000 gpurun_out/r02_bench_v1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_v1.json"))
for k in ("value", "ms_per_step", "e2e", "roofline", "cpu_baseline", "parity", "precision", "group_norm"):
    print(k, json.dumps(d[k])[:1600])
for k, v in d["mesh_512"].items():
    print(k, json.dumps(v)[:2000])
PY
