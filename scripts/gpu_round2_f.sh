#!/bin/bash
python -m pytest tests/test_octree_mc_gpu.py tests/test_fullsize_gpu.py tests/test_postprocess_gpu.py -m gpu -x -q 2>&1 | tail -3
PIFU_MC_CLASSIFY=0 python -m pytest tests/test_octree_mc_gpu.py -m gpu -x -q -k "marching" 2>&1 | tail -2
python scripts/profile_mc.py 512 5
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_mesh512_v9_launches.csv python scripts/profile_mesh.py 512 octree > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_mesh512_v9_launches.csv 40 | grep -E "total|classify|emit|scan_|active_rows"
python scripts/mesh_latency.py > gpurun_out/r02_mesh512_v9.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_mesh512_v9.json"))
for m in ("octree", "dense"):
    print(m, {k: d[m][k] for k in ("latency_ms", "field_ms", "mc_ms", "mesh_d2h_ms")}, d[m]["mc_roofline"]["frac"])
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_octree_mc_gpu.py -m gpu -x -q -k "marching" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head -5
