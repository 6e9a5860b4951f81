"""GPU: the call forms and state the reference API exposes beyond the single-view query
(`PIFuMRNet.py:119-186` multi-crop form, batches, `calc_normal` of both nets, perspective),
checked per item against the CPU oracle and for the reference's tensor shapes / stacking order."""
import pytest
import torch

from helpers import calibrated_problem, oracle_states, orc, syn
from pifu_b200 import config
from test_query_gpu import OCC_TOL, build_nets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def prob():
    torch.set_grad_enabled(False)
    return calibrated_problem()[0]


def _variants(prob, B1, B2):
    """Per-image coarse maps [B1, ...] and per-crop fine maps [B1 * B2, ...] (deterministic variants)."""
    fc = torch.cat([prob["feat_coarse"].roll(7 * b, dims=3) * (1.0 + 0.05 * b) for b in range(B1)], 0)
    ff = torch.cat([prob["feat_fine"].roll(11 * k, dims=2) * (1.0 - 0.03 * k) for k in range(B1 * B2)], 0)
    return fc, ff


def test_multi_crop_batched_query_matches_oracle_per_item(prob):
    B1, B2, N = 2, 3, 700
    netG, netMR = build_nets(prob)
    fc, ff = _variants(prob, B1, B2)
    netG.im_feat_list = [fc.cuda()]
    netMR.im_feat_list = [ff.cuda()]
    pts = torch.stack([torch.stack([syn.random_points(N, 300 + 10 * b1 + b2)[0] for b2 in range(B2)], 0)
                       for b1 in range(B1)], 0)                                   # [B1, B2, 3, N]
    calib_g = torch.cat([syn.default_calib(), syn.scaled_calib(1.125, (0.0, 0.03125, 0.0))], 0)   # [B1, 4, 4]
    calib_l = torch.stack([torch.cat([syn.scaled_calib(1.0 + 0.25 * b2, (0.0625 * b2, 0.0, 0.0)) for b2 in range(B2)], 0)
                           for _ in range(B1)], 0)                                # [B1, B2, 4, 4]
    netMR.query(pts.cuda(), calib_l.cuda(), calib_g.cuda())
    preds, interm, low = netMR.get_preds().cpu(), netMR.preds_interm.cpu(), netMR.preds_low.cpu()
    # reference shapes: preds = cat over crops of [B1,1,N]; interm / low = cat over crops on dim 1
    assert preds.shape == (B2 * B1, 1, N) and interm.shape == (1, B2 * B1, 1, N) and low.shape == (1, B2 * B1, 1, N)
    oc, of = config.coarse_opt(), config.fine_opt()
    for b2 in range(B2):
        for b1 in range(B1):
            coarse = orc.CoarseState(prob["coarse"], fc[b1:b1 + 1], oc)
            fine = orc.FineState(prob["fine"], ff[b1 * B2 + b2:b1 * B2 + b2 + 1], of, coarse)
            ref, ref_low, _ = orc.query_fine(fine, pts[b1:b1 + 1, b2], calib_l[b1:b1 + 1, b2], calib_g[b1:b1 + 1])
            k = b2 * B1 + b1
            assert (preds[k] - ref[0]).abs().max().item() < OCC_TOL
            assert (interm[0, k] - ref[0]).abs().max().item() < OCC_TOL
            assert (low[0, k] - ref_low[0, 0]).abs().max().item() < OCC_TOL
    # netG keeps the state of its last call (the last crop), like the reference
    assert netG.phi.shape == (B1, 256, N)


def test_coarse_batch_and_calc_normal(prob):
    B, N = 3, 500
    netG, _ = build_nets(prob)
    fc, _ = _variants(prob, B, 1)
    netG.im_feat_list = [fc.cuda()]
    pts = torch.cat([syn.random_points(N, 400 + b) for b in range(B)], 0)         # [B, 3, N]
    calib = torch.cat([syn.default_calib()] * B, 0)
    netG.query(pts.cuda(), calib.cuda())
    out, phi = netG.get_preds().cpu(), netG.phi.cpu()
    assert out.shape == (B, 1, N) and phi.shape == (B, 256, N)
    oc = config.coarse_opt()
    for b in range(B):
        coarse = orc.CoarseState(prob["coarse"], fc[b:b + 1], oc)
        ref, ref_phi = orc.query_coarse(coarse, pts[b:b + 1], calib[b:b + 1])
        assert (out[b] - ref[0]).abs().max().item() < OCC_TOL
        assert (phi[b] - ref_phi[0]).abs().max().item() < 2e-3 * max(1.0, ref_phi.abs().max().item())
    # calc_normal (`PIFuNetwNML.py:181-220`): forward differences with delta = 0.1, no in-bounds mask
    netG.calc_normal(pts.cuda(), calib.cuda(), delta=0.1)
    nml = netG.nml.cpu()
    assert nml.shape == (B, 3, N)
    for b in range(B):
        coarse = orc.CoarseState(prob["coarse"], fc[b:b + 1], oc)
        p4 = [pts[b:b + 1].clone() for _ in range(4)]
        for a in range(3):
            p4[a + 1][:, a, :] += 0.1
        pall = torch.stack(p4, 3).view(1, 3, -1)
        xyz = orc.project(pall, calib[b:b + 1], "orthogonal")
        feat = torch.cat([orc.index(coarse.feat, xyz[:, :2, :]), orc.depth_normalize(xyz, oc.loadSize, oc.z_size)], 1)
        pred = orc.mlp_forward(feat, coarse.sd, coarse.n_layers, oc.mlp_res_layers, coarse.merge)[0].view(1, 1, -1, 4)
        raw = -torch.cat([pred[:, :, :, a + 1] - pred[:, :, :, 0] for a in range(3)], 1)
        ok = raw[0].norm(dim=0) > 5e-3                      # differences well above the 16-bit operand noise
        assert ok.sum() > 20
        cos = (nml[b] * torch.nn.functional.normalize(raw[0], dim=0, eps=1e-8)).sum(0)
        assert (cos[ok] > 0.98).all()


def test_perspective_two_level(prob):
    """Perspective projection (any mode string other than 'orthogonal', `BasePIFuNet.py:79`) through both
    levels; the typo default 'otthogonal' of `PIFuMRNet.py:22` therefore selects it too."""
    from pifu_b200 import PIFuMRNet, PIFuNetwNML
    netG = PIFuNetwNML(config.coarse_opt(), "perspective")
    netMR = PIFuMRNet(config.fine_opt(), netG)            # default projection_mode: the reference's typo
    assert netMR.is_perspective and netG.is_perspective
    netG.mlp.load_state_dict(prob["coarse"])
    netMR.mlp.load_state_dict(prob["fine"])
    netMR.cuda().eval()
    netG.im_feat_list = [prob["feat_coarse"].cuda()]
    netMR.im_feat_list = [prob["feat_fine"].cuda()]
    _, fine = oracle_states(prob, mode="perspective")
    pts = syn.random_points(3000, 17).clone()
    pts[:, 2, :] = pts[:, 2, :] * 0.25 + 2.0
    calib = syn.default_calib()
    ref, ref_low, _ = orc.query_fine(fine, pts, calib)
    netMR.query(pts.cuda(), calib.cuda())
    assert (netMR.get_preds().cpu() - ref).abs().max().item() < OCC_TOL
    assert (netMR.preds_low.cpu() - ref_low).abs().max().item() < OCC_TOL


def test_resnapshot_after_load_state_dict_and_new_features(prob):
    """The engine re-snapshots weights after load_state_dict / in-place edits and feature maps after a new
    filter() result, like reading the live nn.Module would."""
    netG, netMR = build_nets(prob)
    pts = syn.random_points(1000, 9).cuda()
    calib = syn.default_calib().cuda()
    netMR.query(pts, calib)
    a = netMR.get_preds().clone()
    sd = {k: v.clone() for k, v in prob["fine"].items()}
    sd["filters.3.bias"] = sd["filters.3.bias"] + 0.5
    netMR.mlp.load_state_dict(sd)
    netMR.query(pts, calib)
    b = netMR.get_preds().clone()
    assert (b - a).abs().max().item() > 1e-2
    with torch.no_grad():
        netMR.mlp.filters[3].bias.sub_(0.5)                # in-place edit bumps the tensor version
    netMR.query(pts, calib)
    assert torch.allclose(netMR.get_preds(), a, atol=1e-6)
    netMR.im_feat_list = [prob["feat_fine"].cuda() * 1.5]
    netMR.query(pts, calib)
    assert (netMR.get_preds() - a).abs().max().item() > 1e-3
