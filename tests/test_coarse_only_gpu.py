"""BASELINE configs[0] on the GPU: the coarse net alone (`PIFuNetwNML.query`, `PIFuNetwNML.py:99-141`) under
`mesh_util.reconstruction` (`mesh_util.py:67-74` calls net.query / net.get_preds on whatever net it is given) -
dense `eval_grid` with levels = 1, the octree, and the mesh, against the CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import oracle_states, orc, syn
from test_chain_gpu import lattice_points
from test_query_gpu import OCC_TOL, sign_agreement

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def coarse():
    """Coarse net whose own last layer (Conv1d 385 -> 1, `MLP.py:72-73`) is calibrated so that its field has an
    iso-surface: the same recipe as the fine net's, applied to filters.4."""
    from pifu_b200 import PIFuNetwNML, config
    torch.set_grad_enabled(False)
    prob = syn.make_problem(bias_std=0.01)
    cst, _ = oracle_states(prob)
    pilot = syn.random_points(20000, syn.SEED_PILOT, -1.0, 1.0)
    p = orc.query_coarse(cst, pilot, syn.default_calib())[0].numpy()
    syn.calibrate_last_layer(prob["coarse"], 4, p)
    syn.saturate(prob["coarse"], 4, gain=4.0)
    cst, _ = oracle_states(prob)
    netG = PIFuNetwNML(config.coarse_opt(), "orthogonal")
    netG.mlp.load_state_dict(prob["coarse"])
    netG.to("cuda").eval()
    netG.im_feat_list = [prob["feat_coarse"].cuda()]
    eng = netG._engine_for(torch.zeros(1, device="cuda"))
    eng.sync_features(0, netG.im_feat_list[-1])
    yield prob, cst, netG, eng
    eng.set_precision("fast")


def test_coarse_eval_grid_vs_oracle(coarse):
    """pifu_eval_grid with levels = 1 on a 128^3 lattice (configs[0]'s size), every 7th point against the oracle."""
    _, cst, netG, eng = coarse
    calib = syn.default_calib()
    R = (128, 128, 128)
    ids = np.arange(0, 128 ** 3, 7)
    ref = orc.query_coarse(cst, lattice_points(R, calib, ids), calib)[0].numpy().ravel()
    eng.set_precision("hybrid")
    try:
        out = eng.eval_grid(1, 128, calib[0]).cpu().numpy()[ids]
    finally:
        eng.set_precision("fast")
    err = np.abs(out - ref).max()
    print("coarse-only 128^3: max |err| %.3e, sign agreement %.6f, occupied %.4f" % (err, sign_agreement(out, ref), (ref > 0.5).mean()))
    assert err < OCC_TOL
    assert sign_agreement(out, ref) >= 0.9999
    assert np.array_equal(out == 0, ref == 0)


def test_coarse_query_ragged_and_masks(coarse):
    _, cst, netG, _ = coarse
    calib = syn.default_calib()
    for n in (1, 129, 5000):
        pts = syn.random_points(n, 300 + n)
        ref, ref_phi = orc.query_coarse(cst, pts, calib)
        netG.query(pts.cuda(), calib.cuda())
        out = netG.get_preds().cpu()
        assert torch.equal(out == 0, ref == 0)
        assert (netG.phi.cpu() - ref_phi).abs().max().item() < 2e-3 * max(1.0, ref_phi.abs().max().item())
        # fast arithmetic on a x4-saturated coarse field: the same logit-domain bound as DESIGN §4
        assert (out - ref).abs().max().item() < 4e-3


@pytest.mark.parametrize("use_octree", [False, True])
def test_coarse_reconstruction_vs_oracle(coarse, use_octree):
    """`reconstruction(netG, ...)` end to end: the field equals the oracle's within the gates, and the mesh is the
    CPU marching cubes of that same field, transformed and flipped as `mesh_util.py:87-92` does."""
    from pifu_b200 import mesh_util
    from oracle import mc_oracle
    _, cst, netG, eng = coarse
    calib = syn.default_calib()
    res = 128 if use_octree else 64          # 128: two octree levels (strides 2, 1)
    eng.set_precision("hybrid")
    try:
        mesh = mesh_util.reconstruction(netG, "cuda", calib.cuda(), res, None, None, thresh=0.5, use_octree=use_octree)
        field = mesh_util.eval_field_device(netG, torch.device("cuda"), calib.cuda(), res, use_octree).cpu().numpy()
    finally:
        eng.set_precision("fast")
    coords, mat, calib_inv = orc.lattice_coords(res, calib)
    ef = orc.make_eval_func(lambda p, c: orc.query_coarse(cst, p, c), calib)
    ref = orc.eval_grid_octree(coords, ef, num_samples=10000) if use_octree else orc.eval_grid(coords, ef, 10000)
    assert np.abs(field - ref).max() < (0.03 if use_octree else OCC_TOL)      # octree: midpoint fills amplify nothing, but a
    assert sign_agreement(field, ref) >= 0.9999                              # flipped skip decision moves a whole cell
    assert mesh != -1
    verts, faces, normals, values = mesh
    rv, rf, rn, rvals, _ = mc_oracle.marching_cubes(field.astype(np.float32), 0.5)
    trans = calib_inv @ mat
    assert np.array_equal(faces, rf[:, ::-1])                                # det < 0 for the readData calib
    assert np.abs(verts - (rv @ trans[:3, :3].T + trans[:3, 3])).max() < 1e-5
    assert verts.dtype == np.float64 and faces.dtype == np.int32
