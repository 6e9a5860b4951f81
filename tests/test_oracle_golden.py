"""Pin the CPU oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/make_golden.py in the authoring container).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from helpers import calibrated_problem, golden, oracle_states, orc, syn

TOL = 2e-6          # same library kernels, same op order: only threading/blocking noise


@pytest.fixture(scope="module")
def setup():
    torch.set_grad_enabled(False)
    prob, pilot = calibrated_problem()
    return prob, pilot, golden("query_none.npz")


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_pilot_matches_reference(setup):
    _, pilot, g = setup
    assert np.abs(pilot[0, 0, :4096] - g["pilot_preds"]).max() < TOL


def test_coarse_query(setup):
    prob, _, g = setup
    coarse, _ = oracle_states(prob)
    pts = syn.random_points(2048)
    pred, phi = orc.query_coarse(coarse, pts, syn.default_calib())
    assert np.abs(pred.numpy() - g["coarse_preds"]).max() < TOL
    assert np.abs(phi[:, :, :128].numpy() - g["coarse_phi_128"]).max() < 2e-5
    # masked points are exactly zero on both sides
    assert np.array_equal(pred.numpy() == 0, g["coarse_preds"] == 0)


def test_mr_query_variants(setup):
    prob, _, g = setup
    _, fine = oracle_states(prob)
    pts = syn.random_points(2048)
    calib = syn.default_calib()
    pred, low, _ = orc.query_fine(fine, pts, calib)
    assert np.abs(pred.numpy() - g["mr_preds"]).max() < TOL
    assert np.abs(low.numpy() - g["mr_preds_low"]).max() < TOL
    assert np.abs(pred.numpy()[None] - g["mr_preds_interm"]).max() < TOL
    calib2 = syn.scaled_calib()
    assert np.abs(orc.query_fine(fine, pts, calib2)[0].numpy() - g["mr_preds_scaled_calib"]).max() < TOL
    assert np.abs(orc.query_fine(fine, pts, calib2, calib)[0].numpy() - g["mr_preds_local_global"]).max() < TOL
    frac = float((g["mr_preds"] > 0.5).mean())
    assert 0.005 < frac < 0.1, frac      # calibrated field has an iso-surface


def test_calc_normal(setup):
    prob, _, g = setup
    _, fine = oracle_states(prob)
    pts = syn.random_points(2048)[:, :, :256]
    calib = syn.default_calib()
    n, raw = orc.calc_normal_fine(fine, pts, calib, calib, delta=0.001, return_raw=True)
    n, raw = n.numpy(), raw.numpy()
    gn = g["mr_nmls_256"]
    # finite differences of a saturating sigmoid: compare where the difference is well
    # above fp32 noise (elsewhere the direction is rounding noise in both implementations)
    ok = np.linalg.norm(raw, axis=1) > 1e-4
    assert ok.sum() > 20
    assert ((n * gn).sum(1)[ok] > 0.9999).all()


def test_perspective(setup):
    prob, _, g = setup
    coarse, _ = oracle_states(prob, mode="perspective")
    pts = syn.random_points(2048).clone()
    pts[:, 2, :] = pts[:, 2, :] * 0.25 + 2.0
    pred, _ = orc.query_coarse(coarse, pts, syn.default_calib())
    assert np.abs(pred.numpy() - g["coarse_preds_perspective"]).max() < TOL


def test_group_norm(setup):
    prob, _, _ = setup
    g = golden("query_group.npz")
    _, fine = oracle_states(prob, "group")
    pts = syn.random_points(2048)
    pred, low, _ = orc.query_fine(fine, pts, syn.default_calib())
    assert np.abs(pred.numpy() - g["mr_preds"]).max() < 2e-5
    assert np.abs(low.numpy() - g["mr_preds_low"]).max() < 2e-5


def test_index_closed_form(setup):
    prob, _, _ = setup
    uv = syn.random_points(4096, 11, -1.05, 1.05)[:, :2]
    a = orc.index(prob["feat_fine"], uv).numpy()
    b = orc.index_closed_form(prob["feat_fine"], uv).numpy()
    assert np.abs(a - b).max() < 5e-6


def test_lattice_and_dense_grid(setup):
    prob, _, g = setup
    _, fine = oracle_states(prob)
    calib = syn.default_calib()
    coords, mat, _ = orc.lattice_coords(16, calib)
    assert np.array_equal(sha(coords), g["grid16_coords_sha"])
    assert np.array_equal(mat, g["grid16_mat"])
    ef = orc.make_eval_func(lambda p, c: orc.query_fine(fine, p, c)[0], calib)
    f = orc.eval_grid(coords, ef, 1000)
    assert f.dtype == np.float64 and np.abs(f - g["mr_grid16"]).max() < TOL


def test_octree_with_net(setup):
    _, _, g = setup
    prob, _ = calibrated_problem(saturated=True)
    _, fine = oracle_states(prob)
    calib = syn.default_calib()
    coords, _, _ = orc.lattice_coords(64, calib)
    ef = orc.make_eval_func(lambda p, c: orc.query_fine(fine, p, c)[0], calib)
    f = orc.eval_grid_octree(coords, ef, init_resolution=16, num_samples=100000)
    ref = g["mr_octree64_init16"]
    assert np.abs(f - ref).max() < 1e-5
    assert np.all(f[-1] == 0) and np.all(f[:, -1] == 0) and np.all(f[:, :, -1] == 0)


def _ellipsoid(points):
    x, y, z = points[0], points[1], points[2]
    r = np.sqrt((x / 0.35) ** 2 + (y / 0.8) ** 2 + (z / 0.3) ** 2)
    return np.clip(0.5 + 2.0 * (1.0 - r), 0.0, 1.0).astype(np.float32)


def _ripple(points):
    x, y, z = points[0], points[1], points[2]
    v = 0.5 + 0.6 * (x * y - z * z) + 0.25 * (x * x * x - y * z)
    return np.clip(v, 0.0, 1.0).astype(np.float32)


ANALYTIC = {"ellipsoid": _ellipsoid, "ripple": _ripple}


@pytest.mark.parametrize("name", ["ellipsoid", "ripple"])
@pytest.mark.parametrize("res,init", [(64, 8), (64, 16), (96, 12), (128, 32), (128, 64)])
def test_octree_analytic_bit_exact(name, res, init):
    """Octree bookkeeping is integer/index work: the float64 field must be bit-identical
    to the reference loop's (sha256 of the bytes), and so must the evaluated-point count."""
    g = golden("octree_analytic.npz")
    coords, _, _ = orc.lattice_coords(res, torch.eye(4)[None])
    stats = []
    f = orc.eval_grid_octree(coords, ANALYTIC[name], init_resolution=init,
                             num_samples=50000, stats=stats)
    key = "%s_%d_%d" % (name, res, init)
    assert sum(n for _, n in stats) == int(g[key + "_evaluated"])
    assert np.array_equal(f[::7, ::5, ::3], g[key + "_probe"])
    assert np.array_equal(sha(f), g[key + "_sha"])
