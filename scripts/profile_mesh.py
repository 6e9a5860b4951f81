"""Driver for ncu: one 512^3 (or argv[1]^3) reconstruction, octree or dense field + marching cubes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                              # noqa: E402
from pifu_b200 import mesh_util           # noqa: E402

torch.set_grad_enabled(False)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
octree = (sys.argv[2] if len(sys.argv) > 2 else "octree") == "octree"
dev = torch.device("cuda", 0)
netG, netMR, eng, calib = bench.build_mesh_problem(dev)
torch.cuda.synchronize()
print("PROFILE_BEGIN launches", eng.launch_count())
mesh = mesh_util.reconstruction(netMR, dev, calib.to(dev), res, None, None, use_octree=octree)
torch.cuda.synchronize()
print("ok", -1 if mesh == -1 else (len(mesh[0]), len(mesh[1])), "launches", eng.launch_count())
