"""512^3 reconstruction latency breakdown on one GPU: dense / octree field, marching cubes,
mesh transfer (what bench.py reports as `mesh_512`)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                              # noqa: E402

torch.set_grad_enabled(False)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
netG, netMR, eng, calib = bench.build_mesh_problem(dev)
out = bench.mesh_latency(netMR, eng, calib, dev, res, reps)
print(json.dumps(out))
