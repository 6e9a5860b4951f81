"""Small driver for ncu: a few chunks of the dense-lattice hot path (same kernels as bench.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pifu_b200 import PIFuMRNet, PIFuNetwNML, config, synthetic as syn   # noqa: E402

torch.set_grad_enabled(False)
prob = syn.make_problem()
netG = PIFuNetwNML(config.coarse_opt(), "orthogonal")
netMR = PIFuMRNet(config.fine_opt(), netG, "orthogonal")
netG.mlp.load_state_dict(prob["coarse"])
netMR.mlp.load_state_dict(prob["fine"])
netMR.cuda().eval()
netG.im_feat_list = [prob["feat_coarse"].cuda()]
netMR.im_feat_list = [prob["feat_fine"].cuda()]
eng = netMR._engine_for(torch.zeros(1, device="cuda"))
eng.sync_features(0, netG.im_feat_list[-1])
eng.sync_features(1, netMR.im_feat_list[-1])
calib = syn.default_calib()
res = 256
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 37888
mode = sys.argv[2] if len(sys.argv) > 2 else "dense"
if mode == "dense":
    out = eng.eval_grid(2, res, calib[0], id_begin=res ** 3 // 2, id_end=res ** 3 // 2 + n)
else:
    # octree + marching cubes at 128^3 on the saturated field
    out = eng.eval_grid_octree(2, 128, calib[0], want64=False, want32=True)[1]
    eng.marching_cubes(torch.sigmoid((out - 0.5) * 50), 0.5)
torch.cuda.synchronize()
print("ok", float(out.float().mean()))
