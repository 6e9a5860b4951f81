#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoders_gpu.py tests/test_gen_mesh_flow_gpu.py -m gpu -q --no-header -x 2>&1 | tail -8
timeout 600 python - <<'PY' 2>&1 | tail -12
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
torch.set_grad_enabled(False)
for fused in ("1", "0"):
    os.environ["PIFU_FUSED_BN_RELU"] = fused
    out = bench.encoder_leg(torch.device("cuda", 0), frames=8)
    print("fused", fused, {k: round(v["filter_global_plus_local_ms"], 2) for k, v in out["modes"].items()}, out["frames_256_octree"]["frames_per_s"])
PY
