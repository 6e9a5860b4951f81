"""torchrun, NCCL, N GPUs: sharded reconstruction (dense and octree) against the single-GPU result
of rank 0 on the same net; prints per-mode latency.  Usage:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_gpu_check.py [res]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                              # noqa: E402
from pifu_b200 import mesh_util           # noqa: E402

torch.set_grad_enabled(False)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
import datetime
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
rank, world = dist.get_rank(), dist.get_world_size()
netG, netMR, eng, calib = bench.build_mesh_problem(dev)
cal = calib.to(dev)
solo = None                              # a one-rank group = the single-GPU computation on rank 0
for r in range(world):
    g = dist.new_group([r])
    if r == rank:
        solo = g
ok = True
for octree in (True, False):
    ref = None
    if rank == 0:
        field = mesh_util.eval_field_device(netMR, dev, cal, res, octree, group=solo)
        v, f, n, val = eng.marching_cubes(field, 0.5)
        ref = (v.cpu().numpy(), f.cpu().numpy())
        del field
    best = None
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        mesh = mesh_util.reconstruction(netMR, dev, cal, res, None, None, use_octree=octree)
        torch.cuda.synchronize(); dist.barrier()
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    if rank == 0:
        mat = np.eye(4); mat[0, 0] = mat[1, 1] = mat[2, 2] = 2.0 / res; mat[:3, 3] = -1
        trans = np.linalg.inv(cal[0].cpu().numpy()) @ mat
        rv = (trans[:3, :3] @ ref[0].T + trans[:3, 3:4]).T
        same = np.array_equal(mesh[1], ref[1][:, ::-1]) and np.abs(mesh[0] - rv).max() < 1e-12
        ok = ok and same
        print("world %d res %d %s: %d verts %d faces, identical to single-GPU mesh: %s, latency %.1f ms"
              % (world, res, "octree" if octree else "dense", len(mesh[0]), len(mesh[1]), same, best), flush=True)
    else:
        assert mesh is None
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
