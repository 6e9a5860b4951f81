"""GPU parity of the fused query path (gather + tcgen05 MLP) against the CPU oracle, through the
reference-shaped Python API which calls the C ABI.  Tolerances are north_star's: occupancy
within 1e-3 absolute, >= 99.99 % sign agreement at the 0.5 iso-level."""
import numpy as np
import pytest
import torch

from helpers import calibrated_problem, golden, oracle_states, orc, syn

pytestmark = pytest.mark.gpu
OCC_TOL = 1e-3


def build_nets(prob, mode="orthogonal", device="cuda"):
    from pifu_b200 import PIFuMRNet, PIFuNetwNML, config
    netG = PIFuNetwNML(config.coarse_opt(), mode)
    netMR = PIFuMRNet(config.fine_opt(), netG, mode)
    netG.mlp.load_state_dict(prob["coarse"])
    netMR.mlp.load_state_dict(prob["fine"])
    netMR.to(device).eval()
    netG.im_feat_list = [prob["feat_coarse"].to(device)]
    netMR.im_feat_list = [prob["feat_fine"].to(device)]
    return netG, netMR


@pytest.fixture(scope="module")
def setup():
    torch.set_grad_enabled(False)
    prob, _ = calibrated_problem()
    return prob


def sign_agreement(a, b, level=0.5):
    return float(((a > level) == (b > level)).mean())


IMPL = {"tcgen05": 0, "simt": 1, "tc1": 2}     # CTA-pair tcgen05 (product), CUDA-core check, 1-CTA tcgen05


@pytest.mark.parametrize("impl", ["simt", "tc1", "tcgen05"])
@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (300, 272, 512), (1000, 832, 256), (4096, 1024, 512),
                                   (40000, 320, 1024)])
def test_layer_kernel_vs_torch(impl, M, K, N):
    """One Conv1d(k=1)+leaky_relu layer; fp16 operands, fp32 accumulate -> compare with fp32 torch
    on the same fp16-rounded operands (tight) and on the fp32 operands (fp16 rounding only)."""
    from pifu_b200 import get_engine
    eng = get_engine("cuda")
    eng.set_gemm_impl(IMPL[impl])
    try:
        g = torch.Generator().manual_seed(M + K + N)
        X = torch.randn(M, K, generator=g)
        W = torch.randn(N, K, generator=g) * 0.05
        b = torch.randn(N, generator=g) * 0.1
        Y = eng.debug_gemm(X, W, b, leaky=True).cpu()
        ref16 = torch.nn.functional.leaky_relu(W.half().float() @ X.half().float().T + b[:, None])
        ref16 = ref16.half().float()               # the kernel stores fp16
        err = (Y - ref16).abs().max().item()
        assert err <= max(1.0, ref16.abs().max().item()) * 2 ** -9, err   # <= 1 fp16 ulp at the top of the range
    finally:
        eng.set_gemm_impl(0)


@pytest.mark.parametrize("impl", ["simt", "tc1", "tcgen05"])
def test_mr_query_parity(setup, impl):
    from pifu_b200 import get_engine
    prob = setup
    _, fine = oracle_states(prob)
    netG, netMR = build_nets(prob)
    get_engine("cuda").set_gemm_impl(IMPL[impl])
    try:
        pts = syn.random_points(20000 if impl == "simt" else 200000, 5)
        calib = syn.default_calib()
        ref, ref_low, ref_phi = orc.query_fine(fine, pts, calib)
        netMR.query(pts.cuda(), calib.cuda())
        out = netMR.get_preds().cpu()
        assert out.shape == ref.shape
        assert (out - ref).abs().max().item() < OCC_TOL
        assert sign_agreement(out.numpy(), ref.numpy()) >= 0.9999
        assert (netMR.preds_low.cpu() - ref_low).abs().max().item() < OCC_TOL
        assert torch.equal(out == 0, ref == 0)         # identical in-bounds masks
        phi = netG.phi.cpu()
        assert (phi - ref_phi).abs().max().item() < 2e-3 * max(1.0, ref_phi.abs().max().item())
    finally:
        get_engine("cuda").set_gemm_impl(0)


def test_golden_points(setup):
    """Same 2048 points the reference itself was run on (tests/golden/query_none.npz)."""
    prob = setup
    g = golden("query_none.npz")
    netG, netMR = build_nets(prob)
    pts = syn.random_points(2048).cuda()
    calib = syn.default_calib().cuda()
    netMR.query(pts, calib)
    assert np.abs(netMR.get_preds().cpu().numpy() - g["mr_preds"]).max() < OCC_TOL
    assert np.abs(netMR.preds_low.cpu().numpy() - g["mr_preds_low"]).max() < OCC_TOL
    assert np.abs(netMR.preds_interm.cpu().numpy() - g["mr_preds_interm"]).max() < OCC_TOL
    netG.query(pts, calib)
    assert np.abs(netG.get_preds().cpu().numpy() - g["coarse_preds"]).max() < OCC_TOL
    assert np.abs(netG.phi[:, :, :128].cpu().numpy() - g["coarse_phi_128"]).max() < 5e-3
    c2 = syn.scaled_calib().cuda()
    netMR.query(pts, c2)
    assert np.abs(netMR.get_preds().cpu().numpy() - g["mr_preds_scaled_calib"]).max() < OCC_TOL
    netMR.query(pts[:, None], c2[:, None], calib)
    assert np.abs(netMR.get_preds().cpu().numpy() - g["mr_preds_local_global"]).max() < OCC_TOL


def test_perspective_golden(setup):
    prob = setup
    g = golden("query_none.npz")
    netG, _ = build_nets(prob, "perspective")
    pts = syn.random_points(2048).clone()
    pts[:, 2, :] = pts[:, 2, :] * 0.25 + 2.0
    netG.query(pts.cuda(), syn.default_calib().cuda())
    assert np.abs(netG.get_preds().cpu().numpy() - g["coarse_preds_perspective"]).max() < OCC_TOL


def test_calc_normal(setup):
    prob = setup
    _, fine = oracle_states(prob)
    _, netMR = build_nets(prob)
    pts = syn.random_points(2048)[:, :, :256]
    calib = syn.default_calib()
    # finite differences amplify the 16-bit operand noise by 1/delta: compare where the
    # occupancy difference is well above it
    ref, raw = orc.calc_normal_fine(fine, pts, calib, calib, delta=0.01, return_raw=True)
    netMR.calc_normal(pts[:, None].cuda(), calib[:, None].cuda(), calib.cuda(), delta=0.01)
    n = netMR.nmls.cpu().numpy()
    assert n.shape == ref.shape
    ok = np.linalg.norm(raw.numpy(), axis=1) > 5e-3
    assert ok.sum() > 10
    assert ((n * ref.numpy()).sum(1)[ok] > 0.98).all()


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000])
def test_ragged_sizes(setup, n):
    prob = setup
    _, fine = oracle_states(prob)
    _, netMR = build_nets(prob)
    pts = syn.random_points(n, 100 + n)
    calib = syn.default_calib()
    ref = orc.query_fine(fine, pts, calib)[0]
    netMR.query(pts.cuda(), calib.cuda())
    assert (netMR.get_preds().cpu() - ref).abs().max().item() < OCC_TOL


def test_chunk_invariance(setup):
    """mlp_norm='none': a point's value does not depend on the chunking."""
    from pifu_b200 import get_engine
    prob = setup
    _, netMR = build_nets(prob)
    pts = syn.random_points(3000, 77).cuda()
    calib = syn.default_calib().cuda()
    eng = get_engine("cuda")
    netMR.query(pts, calib)
    a = netMR.get_preds().clone()
    eng.set_chunk_tiles(3)
    try:
        netMR.query(pts, calib)
        assert torch.equal(a, netMR.get_preds())
    finally:
        eng.set_chunk_tiles(16 * 148)


def test_saturated_field(setup):
    """Octree / marching-cubes field (last layer x8): the logit-domain error of 16-bit operands
    shows as up to ~6e-3 in occupancy at the surface; the sign at 0.5 must still agree."""
    prob, _ = calibrated_problem(saturated=True)
    _, fine = oracle_states(prob)
    _, netMR = build_nets(prob)
    pts = syn.random_points(200000, 9)
    calib = syn.default_calib()
    ref = orc.query_fine(fine, pts, calib)[0]
    netMR.query(pts.cuda(), calib.cuda())
    out = netMR.get_preds().cpu()
    assert (out - ref).abs().max().item() < 8e-3
    assert sign_agreement(out.numpy(), ref.numpy()) >= 0.9999


def test_dense_grid_parity(setup):
    """eval_grid on a 32^3 lattice vs the oracle's float64 lattice + query (mesh_util.py:59-80)."""
    from pifu_b200 import get_engine
    prob = setup
    _, fine = oracle_states(prob)
    _, netMR = build_nets(prob)
    calib = syn.default_calib()
    res = 32
    coords, _, _ = orc.lattice_coords(res, calib)
    ef = orc.make_eval_func(lambda p, c: orc.query_fine(fine, p, c)[0], calib)
    ref = orc.eval_grid(coords, ef, 10000)
    eng = netMR._engine_for(torch.zeros(1, device="cuda"))
    eng.sync_features(0, netMR.netG.im_feat_list[-1])
    eng.sync_features(1, netMR.im_feat_list[-1])
    out = eng.eval_grid(2, res, calib).cpu().numpy().reshape(res, res, res)
    assert np.abs(out - ref).max() < OCC_TOL
    assert sign_agreement(out, ref) >= 0.9999
    g = golden("query_none.npz")["mr_grid16"]
    out16 = eng.eval_grid(2, 16, calib).cpu().numpy().reshape(16, 16, 16)
    assert np.abs(out16 - g).max() < OCC_TOL
