"""Where the time of gen_mesh's vertex-normal loop goes (`reconstruction.py:58-70`): cProfile of the host side + CUDA events."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                              # noqa: E402
from pifu_b200 import mesh_util           # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
netG, netMR, eng, calib = bench.build_mesh_problem(dev)
cal = calib.to(dev)
mesh = mesh_util.reconstruction(netMR, dev, cal, 512, None, None, use_octree=True)
mv = mesh[0]
verts_tensor = torch.from_numpy(mv.T).unsqueeze(0).to(device=dev).float()


def loop():
    color = np.zeros(mv.shape)
    interval = 50000
    for i in range(len(color) // interval + 1):
        left = i * interval
        right = -1 if i == len(color) // interval else (i + 1) * interval
        netMR.calc_normal(verts_tensor[:, None, :, left:right], cal[:, None], cal)
        color[left:right] = (netMR.nmls.detach().cpu().numpy()[0] * 0.5 + 0.5).T
    return color


for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); loop(); torch.cuda.synchronize()
    print("loop %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile()
pr.enable()
loop()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
