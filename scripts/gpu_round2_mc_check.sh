#!/bin/bash
python -m pytest tests/test_octree_mc_gpu.py tests/test_fullsize_gpu.py tests/test_octree_slab_gpu.py -m gpu -x -q 2>&1 | tail -2
PIFU_MC_CLASSIFY=0 python -m pytest tests/test_octree_mc_gpu.py -m gpu -x -q -k "marching" 2>&1 | tail -1
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_octree_mc_gpu.py -m gpu -x -q -k "marching" 2>&1 | grep -E "ERROR SUMMARY|failed|Invalid" | head -5
lat() { python scripts/mesh_latency.py 512 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({m: round(d[m]['mc_ms'],4) for m in ('octree','dense')})"; }
kern() { ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ab_$1.csv "${@:2}" > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/ab_$1.csv 60 | grep -E "classify|emit_rows|emit_vert|emit_faces"; }
python scripts/profile_mc.py 512 8; lat; lat
echo "-- blob"; kern blobc python scripts/profile_mc.py 512 1
echo "-- octree field"; kern octc python scripts/profile_mesh.py 512 octree
