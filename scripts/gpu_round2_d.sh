#!/bin/bash
python -m pytest tests/test_postprocess_gpu.py tests/test_octree_mc_gpu.py tests/test_fullsize_gpu.py tests/test_coarse_only_gpu.py -m gpu -x -q 2>&1 | tail -12
PIFU_MC_CLASSIFY=0 python -m pytest tests/test_octree_mc_gpu.py -m gpu -x -q -k "marching" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_mesh512_v4_launches.csv python scripts/profile_mesh.py 512 octree > gpurun_out/r02_mesh512_v4.log 2>&1
tail -2 gpurun_out/r02_mesh512_v4.log
python scripts/launch_summary.py gpurun_out/r02_mesh512_v4_launches.csv 40 | grep -E "total|classify|emit|scan|active|fill|cells|commit|frontier|init_todo|zero_last"
python scripts/mesh_latency.py > gpurun_out/r02_mesh512_v4.json 2>gpurun_out/r02_mesh512_v4.err; tail -3 gpurun_out/r02_mesh512_v4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_mesh512_v4.json"))
for m in ("octree", "dense", "octree_hybrid"):
    print(m, {k: d[m][k] for k in d[m] if k in ("latency_ms", "field_ms", "mc_ms", "mesh_d2h_ms", "verts", "mc_roofline")})
PY
