"""SURVEY §8(f) row 4 on the GPU: vertex colours sampled from an image (`reconstruction.py:110-116`) and `meshcleaning`
(`reconstruction.py:325-344`), each against its CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import orc, syn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["orthogonal", "perspective"])
def test_vertex_colors_from_image(mode):
    """projection + index(image, uv) * 0.5 + 0.5 fused in one kernel vs the oracle's torch-CPU restatement."""
    from pifu_b200 import BasePIFuNet, mesh_util
    torch.manual_seed(3)
    img = torch.rand(2, 3, 96, 128) * 2 - 1
    pts = syn.random_points(5000, 41)[0]                       # [3, n], overshooting [-1, 1]: zero padding is exercised
    if mode == "perspective":
        pts = pts.clone()
        pts[2] = pts[2] * 0.25 + 2.0
    calib = syn.scaled_calib(1.1)
    net = BasePIFuNet.BasePIFuNet(projection_mode=mode)
    got = mesh_util.vertex_colors_from_image(net, img.cuda(), pts.T.numpy(), calib.cuda())
    xyz = orc.project(pts[None], calib, mode)
    ref = (orc.index(img[:1], xyz[:, :2, :])[0].T * 0.5 + 0.5).numpy()
    assert got.shape == ref.shape == (5000, 3)
    assert np.abs(got - ref).max() < 2e-6


def _two_blobs(n=48):
    g = np.linspace(-1, 1, n)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    big = 0.45 - np.sqrt((x + 0.2) ** 2 + y ** 2 + z ** 2)                     # tall along x
    small = 0.2 - np.sqrt((x - 0.6) ** 2 + (y - 0.5) ** 2 + (z + 0.4) ** 2)
    wide = 0.3 - np.sqrt(((x - 0.1) / 0.3) ** 2 + (y + 0.55) ** 2 + (z - 0.55) ** 2)   # short along x, longer along y/z
    opened = 0.12 - np.sqrt((y + 0.75) ** 2 + (z + 0.75) ** 2) + 0 * x       # a bar along x, cut by the volume border: the longest, not watertight
    return (np.maximum.reduce([big, small, wide, opened]) + 0.5).astype(np.float32)


@pytest.mark.parametrize("only_watertight", [True, False])
def test_clean_mesh_vs_oracle(only_watertight):
    from oracle import mesh_oracle
    from pifu_b200 import get_engine, mesh_util
    eng = get_engine("cuda")
    v, f, _, _ = eng.marching_cubes(torch.from_numpy(_two_blobs()).cuda(), 0.5)
    verts, faces = v.cpu().numpy(), f.cpu().numpy()
    colors = np.random.default_rng(0).random(verts.shape)
    rv, rf, rc = mesh_oracle.largest_component(verts, faces, colors, only_watertight)
    gv, gf, gc = mesh_util.clean_mesh(verts, faces, colors, device="cuda", only_watertight=only_watertight)
    # the flipped, negative-stride face view `reconstruction()` returns (`mesh_util.py:91-92`) must be accepted as well
    fv, ff, _ = mesh_util.clean_mesh(verts, faces[:, ::-1], None, device="cuda", only_watertight=only_watertight)
    assert np.array_equal(ff, rf[:, ::-1]) and np.array_equal(fv, rv)
    assert 0 < len(rv) < len(verts)
    assert (rv[:, 0].max() - rv[:, 0].min() > 40) == (not only_watertight)      # the open bar wins only when open components count
    assert np.array_equal(gv, rv) and np.array_equal(gf, rf) and np.array_equal(gc, rc)
    assert gf.dtype == np.int32 and gv.dtype == np.float64


def test_meshcleaning_file_roundtrip(tmp_path):
    """`meshcleaning(obj_path)` as the reference calls it: file in, file out."""
    from oracle import mesh_oracle
    from pifu_b200 import get_engine, mesh_util
    eng = get_engine("cuda")
    v, f, _, _ = eng.marching_cubes(torch.from_numpy(_two_blobs(32)).cuda(), 0.5)
    verts, faces = v.cpu().numpy(), f.cpu().numpy()
    colors = np.round(np.random.default_rng(1).random(verts.shape), 4)
    path = str(tmp_path / "m.obj")
    mesh_util.save_obj_mesh_with_color(path, verts, faces, colors)
    lv, lf, lc = mesh_util.load_obj_mesh_with_color(path)
    assert np.array_equal(lf, faces) and np.abs(lv - verts).max() < 1e-4
    mesh_util.meshcleaning(path, device="cuda")
    cv, cf, cc = mesh_util.load_obj_mesh_with_color(path)
    rv, rf, rc = mesh_oracle.largest_component(lv, lf, lc)
    assert np.array_equal(cf, rf) and np.abs(cv - rv).max() < 1e-4 and np.abs(cc - rc).max() < 1e-4


def test_clean_mesh_without_candidate_raises():
    from pifu_b200 import _lib, mesh_util
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)
    faces = np.array([[0, 1, 2]], dtype=np.int32)
    with pytest.raises(_lib.PifuError):
        mesh_util.clean_mesh(verts, faces, device="cuda")          # one open triangle: trimesh's cc would be empty
