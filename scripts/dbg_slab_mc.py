"""Debug driver: the slab marching-cubes calls rank r of W makes in dist.sharded_mesh (octree field first, dense second,
same hint key), on one GPU without NCCL."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pifu_b200 import mesh_util, dist as pdist

torch.set_grad_enabled(False)
res, W = 512, 8
dev = torch.device("cuda", 0)
netG, netMR, eng, calib = bench.build_mesh_problem(dev)
cal = calib.to(dev)
oct_field = mesh_util.eval_field_device(netMR, dev, cal, res, True)
dense_field = mesh_util.eval_field_device(netMR, dev, cal, res, False)
for r in range(W):
    pb, pe = r * res // W, (r + 1) * res // W
    lo, hi = max(pb - 1, 0), min(pe + 2, res)
    cells_end = min(pe, res - 1)
    for name, field in (("octree", oct_field), ("dense", dense_field)):
        sub = field[lo:hi].clone()
        v, f, n, val, cnt = eng.marching_cubes_slab_async(sub, 0.5, lo, res, cells_end - lo, pb > 0)
        tv, tf, ng = (int(x) for x in cnt.tolist())
        over = tv > v.shape[0] or tf > f.shape[0]
        if over:
            v, f, n, val, cnt = eng.marching_cubes_slab_async(sub, 0.5, lo, res, cells_end - lo, pb > 0, cap=(tv, tf))
        eng.marching_cubes_note_counts(tv, tf)
        frag = pdist._pack_fragment(v[ng:tv], f[:tf] + 7, n[ng:tv], val[ng:tv])
        torch.cuda.synchronize()
        print("rank %d %s: verts %d faces %d ghost %d cap (%d, %d) overflow %s frag %d" % (r, name, tv, tf, ng, v.shape[0], f.shape[0], over, frag.numel()), flush=True)
