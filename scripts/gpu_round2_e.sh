#!/bin/bash
python scripts/profile_mc.py 512 5
PIFU_MC_CLASSIFY=0 python scripts/profile_mc.py 512 5
ncu --set full --clock-control none --import-source on -k regex:classify_warp_kernel -c 1 -o gpurun_out/r02_mc_classify_warp python scripts/profile_mc.py 512 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:emit_rows_kernel -c 1 -o gpurun_out/r02_mc_emit_rows python scripts/profile_mc.py 512 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
python -m pytest tests/test_precision_gpu.py -m gpu -x -q -s 2>&1 | tail -22
python - <<'PY'
import sys, time, torch
sys.path.insert(0, ".")
import bench
from pifu_b200 import mesh_util
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
netG, netMR, eng, calib = bench.build_mesh_problem(dev)
cal = calib.to(dev)
out = torch.empty(256 ** 3, device=dev)
for mode in ("fast", "hybrid", "split"):
    eng.set_precision(mode)
    eng.eval_grid(2, 256, calib[0], out=out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        eng.eval_grid(2, 256, calib[0], out=out)
    ev[1].record(); torch.cuda.synchronize()
    print("dense 256^3 %s: %.2f ms" % (mode, ev[0].elapsed_time(ev[1]) / 3))
for mode in ("fast", "hybrid"):
    eng.set_precision(mode)
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mesh_util.reconstruction(netMR, dev, cal, 512, None, None, use_octree=True)
        torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) * 1e3)
    print("octree 512^3 %s: %.2f ms" % (mode, best))
eng.set_precision("fast")
PY
