"""Host-side phase timing of mesh_util.reconstruction (octree, 512^3): where the wall clock goes beyond the kernels."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                              # noqa: E402
from pifu_b200 import mesh_util           # noqa: E402

torch.set_grad_enabled(False)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
netG, netMR, eng, calib = bench.build_mesh_problem(dev)
cal = calib.to(dev)


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(4):
    t = [T()]
    e, levels = mesh_util._prepare_native(netMR, dev); t.append(T())
    _, sdf32, ev = e.eval_grid_octree(levels, res, cal[0], 64, 0.05, want64=False, want32=True); t.append(T())
    verts, faces, normals, values = e.marching_cubes(sdf32, 0.5); t.append(T())
    calib_inv = np.linalg.inv(cal[0].detach().cpu().numpy())
    mat = np.eye(4); mat[0, 0] = mat[1, 1] = mat[2, 2] = 2.0 / res; mat[0:3, 3] = -1.0
    trans = np.matmul(calib_inv, mat)
    tt = torch.from_numpy(trans).to(dev)
    v2 = verts @ tt[:3, :3].T + tt[:3, 3]; t.append(T())
    out = (v2.cpu().numpy(), faces.cpu().numpy(), normals.cpu().numpy(), values.cpu().numpy()); t.append(T())
    t0 = T()
    mesh = mesh_util.reconstruction(netMR, dev, cal, res, None, None, use_octree=True)
    t1 = T()
    print("rep %d: prepare %.2f  octree field %.2f  mc %.2f  transform %.2f  d2h %.2f  | sum %.2f  reconstruction() %.2f ms" % (
        (rep,) + tuple((t[i + 1] - t[i]) * 1e3 for i in range(5)) + ((t[5] - t[0]) * 1e3, (t1 - t0) * 1e3)))
# the octree driver level by level (stepwise ABI): frontier / evaluation / commit
for rep in range(2):
    e.octree_begin(res, 64, 0.05)
    line = []
    while True:
        a = T()
        step, ids = e.octree_frontier()
        b = T()
        if step == 0:
            break
        vals = e.eval_lattice_ids(2, res, ids, cal[0])
        c = T()
        e.octree_commit(vals)
        d = T()
        line.append("step %d n %d: frontier %.2f eval %.2f commit %.2f" % (step, ids.numel(), (b - a) * 1e3, (c - b) * 1e3, (d - c) * 1e3))
    a = T(); e.octree_export(want64=False, want32=True); b = T()
    print(" | ".join(line), "| export %.2f" % ((b - a) * 1e3))
