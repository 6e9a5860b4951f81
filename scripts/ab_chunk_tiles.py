"""A/B of the rows per launch of the run-list chain (pifu_set_chunk_tiles) on the 512^3 octree reconstruction."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                              # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
netG, netMR, eng, calib = bench.build_mesh_problem(dev)
for tiles in (2368, 1184, 4736, 9472, 2368):
    eng.set_chunk_tiles(tiles)
    out = bench.mesh_latency(netMR, eng, calib, dev, 512, 4)
    o = out["octree"]
    print("chunk_tiles %5d: octree latency %.2f ms, field %.2f ms, mc %.3f ms" % (tiles, o["latency_ms"], o["field_ms"], o["mc_ms"]), flush=True)
