"""GPU parity of the normalised MLP stacks (`MLP.py:36-41,66-69`; mlp_norm='group' is the
reference's default, options.py:95) against the CPU oracle, through the reference-shaped API.
The statistics run over the points of one call, so the tests also pin that coupling: the same
point evaluated in a different call gives a different value, exactly as in the reference."""
import numpy as np
import pytest
import torch

from helpers import calibrated_problem, golden, orc, syn
from pifu_b200 import config
from test_query_gpu import OCC_TOL, sign_agreement

pytestmark = pytest.mark.gpu


def _norm_params(dims, seed, batch):
    rng = np.random.default_rng(seed)
    sd = {}
    for i in range(len(dims) - 2):
        c = dims[i + 1]
        sd["norms.%d.weight" % i] = torch.from_numpy((1.0 + 0.2 * rng.standard_normal(c)).astype(np.float32))
        sd["norms.%d.bias" % i] = torch.from_numpy((0.1 * rng.standard_normal(c)).astype(np.float32))
        if batch:
            sd["norms.%d.running_mean" % i] = torch.from_numpy((0.05 * rng.standard_normal(c)).astype(np.float32))
            sd["norms.%d.running_var" % i] = torch.from_numpy((0.5 + rng.uniform(0, 1, c)).astype(np.float32))
    return sd


def _build(norm, training=False, seed=11, calibrate=True):
    """Nets of this package + oracle states with the same weights and seeded affine parameters."""
    from pifu_b200 import PIFuMRNet, PIFuNetwNML
    torch.set_grad_enabled(False)
    prob, _ = calibrated_problem()
    kind = "group" if norm == "group" else "batch"
    oc, of = config.coarse_opt(mlp_norm=kind), config.fine_opt(mlp_norm=kind)
    sdc = dict(prob["coarse"], **_norm_params(oc.mlp_dim, seed, kind == "batch"))
    sdf = dict(prob["fine"], **_norm_params(of.mlp_dim, seed + 1, kind == "batch"))
    onorm = "group" if norm == "group" else ("batch_train" if training else "batch_eval")
    ooc, oof = config.coarse_opt(mlp_norm=onorm), config.fine_opt(mlp_norm=onorm)
    coarse = orc.CoarseState(sdc, prob["feat_coarse"], ooc)
    fine = orc.FineState(sdf, prob["feat_fine"], oof, coarse)
    if calibrate:
        # normalisation changes the logit scale: re-calibrate the last fine conv (shared by both
        # paths) so the gate field has O(1) logits again (synthetic.calibrate_last_layer)
        pilot = syn.random_points(20000, syn.SEED_PILOT, -1.0, 1.0)
        p = orc.query_fine(fine, pilot, syn.default_calib())[0].numpy()
        syn.calibrate_last_layer(sdf, 3, p)
    netG = PIFuNetwNML(oc, "orthogonal")
    netMR = PIFuMRNet(of, netG, "orthogonal")
    netG.mlp.load_state_dict(sdc, strict=False)
    netMR.mlp.load_state_dict(sdf, strict=False)
    netMR.cuda()
    netMR.eval()
    if training:
        netMR.mlp.train()
        netG.mlp.train()
    netG.im_feat_list = [prob["feat_coarse"].cuda()]
    netMR.im_feat_list = [prob["feat_fine"].cuda()]
    return netG, netMR, coarse, fine


@pytest.mark.parametrize("norm,training", [("group", False), ("batch", True), ("batch", False)])
@pytest.mark.parametrize("n", [5000, 20000, 333])
def test_normalised_query_parity(norm, training, n):
    netG, netMR, coarse, fine = _build(norm, training)
    pts = syn.random_points(n, 40 + n)
    calib = syn.default_calib()
    ref, ref_low, ref_phi = orc.query_fine(fine, pts, calib)
    netMR.query(pts.cuda(), calib.cuda())
    out = netMR.get_preds().cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() < OCC_TOL
    assert (netMR.preds_low.cpu() - ref_low).abs().max().item() < OCC_TOL
    assert torch.equal(out == 0, ref == 0)
    phi = netG.phi.cpu()
    assert (phi - ref_phi).abs().max().item() < 4e-3 * max(1.0, ref_phi.abs().max().item())
    if n >= 20000:
        assert sign_agreement(out.numpy(), ref.numpy()) >= 0.999
    # coarse net alone (PIFuNetwNML.query)
    refc, _ = orc.query_coarse(coarse, pts, calib)
    netG.query(pts.cuda(), calib.cuda())
    assert (netG.get_preds().cpu() - refc).abs().max().item() < OCC_TOL


def test_group_norm_couples_the_points_of_a_call():
    """5 000 points evaluated alone or inside a 20 000-point call differ (by ~1e-4 in
    the reference); the device path must follow the reference in both calls, not be chunk-invariant."""
    netG, netMR, coarse, fine = _build("group")
    pts = syn.random_points(20000, 5)
    calib = syn.default_calib()
    netMR.query(pts.cuda(), calib.cuda())
    big = netMR.get_preds().cpu()[..., :5000]
    netMR.query(pts[..., :5000].cuda(), calib.cuda())
    small = netMR.get_preds().cpu()
    ref_big = orc.query_fine(fine, pts, calib)[0][..., :5000]
    ref_small = orc.query_fine(fine, pts[..., :5000], calib)[0]
    assert (ref_big - ref_small).abs().max().item() > 1e-5           # the reference itself is call-dependent
    assert (big - ref_big).abs().max().item() < OCC_TOL and (small - ref_small).abs().max().item() < OCC_TOL
    d_ref, d_out = (ref_big - ref_small).flatten(), (big - small).flatten()
    assert torch.corrcoef(torch.stack([d_ref, d_out]))[0, 1].item() > 0.9


def test_group_norm_golden_reference_values():
    """tests/golden/query_group.npz: outputs of the unmodified reference with mlp_norm='group'."""
    from pifu_b200 import PIFuMRNet, PIFuNetwNML
    torch.set_grad_enabled(False)
    g = golden("query_group.npz")
    prob, _ = calibrated_problem()
    oc, of = config.coarse_opt(mlp_norm="group"), config.fine_opt(mlp_norm="group")
    netG = PIFuNetwNML(oc, "orthogonal")
    netMR = PIFuMRNet(of, netG, "orthogonal")
    netG.mlp.load_state_dict(prob["coarse"], strict=False)       # GroupNorm keeps its default gamma 1 / beta 0
    netMR.mlp.load_state_dict(prob["fine"], strict=False)
    netMR.cuda().eval()
    netG.im_feat_list = [prob["feat_coarse"].cuda()]
    netMR.im_feat_list = [prob["feat_fine"].cuda()]
    pts = syn.random_points(2048).cuda()
    netMR.query(pts, syn.default_calib().cuda())
    # this field is the reference's run on weights calibrated for mlp_norm='none': with GroupNorm the
    # logits are ~10x larger (saturated field), so the 16-bit-operand error shows like in
    # test_saturated_field
    out = netMR.get_preds().cpu().numpy()
    assert np.abs(out - g["mr_preds"]).max() < 8e-3
    assert sign_agreement(out, g["mr_preds"]) >= 0.999
    assert np.abs(netMR.preds_low.cpu().numpy() - g["mr_preds_low"]).max() < OCC_TOL


@pytest.mark.parametrize("use_octree", [False, True])
def test_group_norm_reconstruction_follows_num_samples(use_octree):
    """mesh_util.reconstruction with a normalised MLP cuts the lattice into the reference's
    num_samples chunks: the field must match the oracle's eval_grid / eval_grid_octree run with the
    same chunking (and the mesh is the MC of that field)."""
    from pifu_b200 import mesh_util
    netG, netMR, coarse, fine = _build("group")
    calib = syn.default_calib()
    res, ns = 32, 5000
    coords, _, _ = orc.lattice_coords(res, calib)
    ef = orc.make_eval_func(lambda p, c: orc.query_fine(fine, p, c)[0], calib)
    if use_octree:
        ref = orc.eval_grid_octree(coords, ef, init_resolution=8, num_samples=ns)
        field = mesh_util.eval_field_device(netMR, "cuda", calib.cuda(), res, True, init_resolution=8, num_samples=ns)
    else:
        ref = orc.eval_grid(coords, ef, ns)
        field = mesh_util.eval_field_device(netMR, "cuda", calib.cuda(), res, False, num_samples=ns)
    out = field.cpu().numpy()
    assert np.abs(out - ref).max() < (3e-3 if use_octree else OCC_TOL)   # octree midpoints add two corner errors
    assert sign_agreement(out, ref) >= 0.999
    m = mesh_util.reconstruction(netMR, "cuda", calib.cuda(), res, None, None, use_octree=use_octree, num_samples=ns)
    assert m == -1 or (m[0].shape[1] == 3 and m[1].shape[1] == 3)
