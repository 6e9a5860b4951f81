"""GPU: the call sequence of the reference's `gen_mesh` (`reconstruction.py:25-75`) driven end to end
through this package's drop-in classes: filter_global / filter_local with caller-supplied PyTorch
encoders, reconstruction(use_octree=True, num_samples=5000), the 50 000-vertex calc_normal colour loop
(including its `left:-1` last chunk) and save_obj_mesh_with_color."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import calibrated_problem, syn
from pifu_b200 import PIFuMRNet, PIFuNetwNML, config, mesh_util

pytestmark = pytest.mark.gpu


class FixedFeatures(nn.Module):
    """Stand-in for the reference's hourglass `Filter` (`Filter.py:186-228` returns (outputs, normx)):
    a 1x1 conv on the image plus a fixed band-limited feature map, so the field has a surface."""

    def __init__(self, feat, in_ch):
        super().__init__()
        self.register_buffer("feat", feat)
        self.mix = nn.Conv2d(in_ch, feat.shape[1], 1, bias=False)
        nn.init.normal_(self.mix.weight, 0.0, 0.01)

    def forward(self, x):
        y = torch.nn.functional.interpolate(self.mix(x), size=self.feat.shape[2:], mode="bilinear", align_corners=False)
        out = self.feat + 0.01 * y
        return [out], out[:, :1]


def test_gen_mesh_call_sequence(tmp_path):
    torch.set_grad_enabled(False)
    prob, _ = calibrated_problem(saturated=True)
    cuda = torch.device("cuda")
    netG = PIFuNetwNML(config.coarse_opt(), "orthogonal", image_filter=FixedFeatures(prob["feat_coarse"], 3))
    netMR = PIFuMRNet(config.fine_opt(), netG, "orthogonal", image_filter=FixedFeatures(prob["feat_fine"], 3))
    netG.mlp.load_state_dict(prob["coarse"])
    netMR.mlp.load_state_dict(prob["fine"])
    netMR.to(cuda)
    netG.eval()                                            # `reconstruction.py:288-289`: only netG.eval() ...
    netMR.eval()                                           # (our stand-in encoder has no train-mode behaviour)
    img = syn.synthetic_rgbd(1024, 101)[:, :3]
    data = {"img_512": torch.nn.functional.interpolate(img, size=(512, 512), mode="bilinear", align_corners=False),
            "img": img, "calib": syn.default_calib(), "b_min": np.array([-1, -1, -1]), "b_max": np.array([1, 1, 1])}
    # ---- reconstruction.py:29-34
    image_tensor_global = data["img_512"].to(device=cuda)
    image_tensor = data["img"].to(device=cuda)
    calib_tensor = data["calib"].to(device=cuda)
    netMR.filter_global(image_tensor_global)
    netMR.filter_local(image_tensor[:, None])
    assert netG.im_feat_list[-1].shape == (1, 256, 128, 128) and netMR.im_feat_list[-1].shape == (1, 16, 512, 512)
    # ---- reconstruction.py:56-57
    res = 128
    out = mesh_util.reconstruction(netMR, cuda, calib_tensor, res, data["b_min"], data["b_max"], 0.5,
                                   use_octree=True, num_samples=5000)
    assert out != -1
    verts, faces, _, _ = out
    assert verts.dtype == np.float64 and faces.dtype == np.int32 and len(verts) > 1000
    assert np.abs(verts).max() <= 1.0 + 1e-9               # lattice [-1, 1)^3 mapped back through inv(calib)
    verts_tensor = torch.from_numpy(verts.T).unsqueeze(0).to(device=cuda).float()
    # ---- reconstruction.py:59-70 (interval shrunk so that several chunks and the `left:-1` tail occur)
    color = np.zeros(verts.shape)
    interval = max(len(color) // 3, 1)
    for i in range(len(color) // interval + 1):
        left = i * interval
        right = -1 if i == len(color) // interval else (i + 1) * interval
        netMR.calc_normal(verts_tensor[:, None, :, left:right], calib_tensor[:, None], calib_tensor)
        nml = netMR.nmls.detach().cpu().numpy()[0] * 0.5 + 0.5
        color[left:right] = nml.T
    assert np.isfinite(color).all() and color.min() >= 0.0 and color.max() <= 1.0
    assert (color[:-1].std(0) > 0.01).all()                # real normals, not a constant
    # ---- reconstruction.py:72
    path = str(tmp_path / "mesh.obj")
    mesh_util.save_obj_mesh_with_color(path, verts, faces, color)
    lines = open(path).read().splitlines()
    assert len(lines) == len(verts) + len(faces)
    assert lines[0].startswith("v ") and lines[-1].startswith("f ")
    v0 = [float(x) for x in lines[0].split()[1:]]
    assert np.allclose(v0[:3], verts[0], atol=5e-5) and np.allclose(v0[3:], color[0], atol=5e-5)
    f0 = [int(x) for x in lines[len(verts)].split()[1:]]
    assert f0 == [faces[0][0] + 1, faces[0][2] + 1, faces[0][1] + 1]
