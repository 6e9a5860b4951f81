"""Driver for ncu / timing: marching cubes alone on a synthetic res^3 float32 volume (a wavy blob, ~1 % surface cells),
argv: res [reps].  Prints the event-timed duration per extraction and the achieved algorithmic GB/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pifu_b200 import get_engine           # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
g = torch.linspace(-1, 1, res, device=dev)
x, y, z = torch.meshgrid(g, g, g, indexing="ij")
vol = torch.sigmoid(40 * (0.55 + 0.08 * torch.sin(9 * x) * torch.cos(7 * y) - torch.sqrt((x / 0.5) ** 2 + (y / 0.9) ** 2 + (z / 0.45) ** 2))).float().contiguous()
del x, y, z
eng = get_engine(dev)
v, f, n, val = eng.marching_cubes(vol, 0.5)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
tot = 0.0
for _ in range(reps):
    flush.fill_(0.0)
    ev[0].record()
    v, f, n, val = eng.marching_cubes(vol, 0.5)
    ev[1].record()
    torch.cuda.synchronize()
    tot += ev[0].elapsed_time(ev[1])
ms = tot / reps
nbytes = 4.0 * res ** 3 + v.shape[0] * 40 + f.shape[0] * 12
print("res %d: %d verts %d faces, %.3f ms per extraction, %.0f GB/s algorithmic" % (res, v.shape[0], f.shape[0], ms, nbytes / ms / 1e6))
