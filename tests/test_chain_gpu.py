"""GPU parity of the lattice chain kernel (whole MLP stack of a lattice-column tile in one
kernel) against the CPU oracle and against the per-layer kernels, through the C ABI
(`pifu_eval_grid`).  Tolerances are north_star's: occupancy within 1e-3 absolute, >= 99.99 %
sign agreement at the 0.5 iso-level."""
import numpy as np
import pytest
import torch

from helpers import calibrated_problem, oracle_states, orc, syn
from test_query_gpu import OCC_TOL, build_nets, sign_agreement

pytestmark = pytest.mark.gpu


def lattice_points(R, calib, ids):
    """float64 lattice + calib pre-transform + float32 cast (`mesh_util.py:12-38,59-65,70`)."""
    R0, R1, R2 = R
    k, j, i = ids % R2, (ids // R2) % R1, ids // (R1 * R2)
    c = np.stack([(2.0 / R0) * i + (-1.0), (2.0 / R1) * j + (-1.0), (2.0 / R2) * k + (-1.0),
                  np.ones(len(ids))], 1)
    inv = np.linalg.inv(calib[0].numpy()).astype(np.float64)
    p = (c @ inv.T)[:, :3].T
    return torch.from_numpy(np.ascontiguousarray(p.astype(np.float32)))[None]


@pytest.fixture(scope="module")
def setup():
    torch.set_grad_enabled(False)
    prob, _ = calibrated_problem()
    _, fine = oracle_states(prob)
    netG, netMR = build_nets(prob)
    eng = netMR._engine_for(torch.zeros(1, device="cuda"))
    eng.sync_features(0, netG.im_feat_list[-1])
    eng.sync_features(1, netMR.im_feat_list[-1])
    pts = syn.random_points(256)
    netMR.query(pts.cuda(), syn.default_calib().cuda())        # snapshots both MLPs
    assert eng.chain_ready()
    return prob, fine, netMR, eng


CASES = [
    # lattice, id range, calib            what it exercises
    ((2, 1, 128), None, "default"),     # one CTA pair, one tile each
    ((3, 1, 128), None, "default"),     # odd tile count: the peer CTA of the last pair idles
    ((4, 8, 256), None, "default"),     # two tiles per column
    ((6, 5, 256), (100, 7000), "default"),   # ragged ends go through the per-layer kernels
    ((4, 8, 128), None, "scaled"),      # scale + shift calibration (some columns out of bounds)
    ((24, 32, 128), None, "default"),   # 768 tiles: several tiles per CTA, accumulator halves swap
]


@pytest.mark.parametrize("R,rng,calib_name", CASES, ids=["2x1x128", "3x1x128", "4x8x256", "6x5x256-ragged", "4x8x128-scaled", "24x32x128"])
def test_chain_vs_oracle(setup, R, rng, calib_name):
    prob, fine, netMR, eng = setup
    calib = syn.default_calib() if calib_name == "default" else syn.scaled_calib()
    total = R[0] * R[1] * R[2]
    a, b = rng if rng else (0, total)
    ids = np.arange(a, b)
    ref = orc.query_fine(fine, lattice_points(R, calib, ids), calib)[0].numpy().ravel()
    eng.set_chain(True)
    l0 = eng.launch_count()
    out = eng.eval_grid(2, R, calib[0], id_begin=a, id_end=b).cpu().numpy()
    assert eng.launch_count() - l0 <= 4 + 14, "chain path not taken"
    assert np.abs(out - ref).max() < OCC_TOL
    assert np.array_equal(out == 0, ref == 0)              # identical in-bounds masks
    if len(ids) >= 50000:
        assert sign_agreement(out, ref) >= 0.9999
    eng.set_chain(False)
    try:
        layer = eng.eval_grid(2, R, calib[0], id_begin=a, id_end=b).cpu().numpy()
    finally:
        eng.set_chain(True)
    assert np.abs(out - layer).max() < OCC_TOL


def test_eval_grid_host_delivery(setup):
    """Engine.eval_grid_host (field to pinned host memory, copies of finished pieces overlapping the next piece's
    kernels) returns exactly what eval_grid + .cpu() returns - whole volume, a ragged sub-range, and a range shorter
    than one piece."""
    _, _, _, eng = setup
    calib = syn.default_calib()
    R = (80, 64, 256)                                   # 5 120 columns: more than one launch (4 736 columns on 148 SMs)
    total = R[0] * R[1] * R[2]
    for a, b, lpp in ((0, total, 1), (256 * 13, total - 256 * 7, 1), (0, 256 * 40, 4)):
        want = eng.eval_grid(2, R, calib[0], id_begin=a, id_end=b).cpu()
        host = torch.full((b - a,), -1.0).pin_memory()
        got = eng.eval_grid_host(2, R, calib[0], host, id_begin=a, id_end=b, launches_per_piece=lpp)
        torch.cuda.synchronize()
        assert got is host and torch.equal(host, want)
    with pytest.raises(ValueError):
        eng.eval_grid_host(2, R, calib[0], torch.empty(5), id_begin=0, id_end=256)


def test_chain_not_taken_when_z_mixes_into_xy(setup):
    """A calibration that rotates z into x makes the samples vary along the column: the dense
    path must fall back to the per-layer kernels and still match the oracle."""
    prob, fine, netMR, eng = setup
    c = syn.default_calib().clone()
    c[0, 0, 0], c[0, 0, 2], c[0, 2, 0], c[0, 2, 2] = 0.8, 0.6, -0.6, 0.8
    R = (4, 4, 128)
    ids = np.arange(R[0] * R[1] * R[2])
    ref = orc.query_fine(fine, lattice_points(R, c, ids), c)[0].numpy().ravel()
    out = eng.eval_grid(2, R, c[0]).cpu().numpy()
    assert np.abs(out - ref).max() < OCC_TOL


def test_chain_saturated_field():
    """Octree / marching-cubes field (last layer x8), 128^2 columns x 128: sign agreement."""
    torch.set_grad_enabled(False)
    prob, _ = calibrated_problem(saturated=True)
    _, fine = oracle_states(prob)
    netG, netMR = build_nets(prob)
    eng = netMR._engine_for(torch.zeros(1, device="cuda"))
    eng.sync_features(0, netG.im_feat_list[-1])
    eng.sync_features(1, netMR.im_feat_list[-1])
    calib = syn.default_calib()
    netMR.query(syn.random_points(256).cuda(), calib.cuda())
    R = (40, 40, 128)
    ids = np.arange(R[0] * R[1] * R[2])
    ref = orc.query_fine(fine, lattice_points(R, calib, ids), calib)[0].numpy().ravel()
    out = eng.eval_grid(2, R, calib[0]).cpu().numpy()
    assert np.abs(out - ref).max() < 8e-3
    assert sign_agreement(out, ref) >= 0.9999


# ---------------------------------------------------------------- run-list form (octree frontiers)
def band_ids(R, width, seed):
    """Sorted lattice ids of a wavy band |k - f(i, j)| < w(i, j): per column one run of 1 .. 2 width points,
    the shape of an octree frontier around a surface (`mesh_util.py:142-149`)."""
    R0, R1, R2 = R
    g = np.random.default_rng(seed)
    i, j = np.meshgrid(np.arange(R0), np.arange(R1), indexing="ij")
    centre = R2 / 2 + 0.3 * R2 * np.sin(0.37 * i + 0.2) * np.cos(0.23 * j)
    w = g.integers(1, width + 1, size=(R0, R1))
    k = np.arange(R2)[None, None, :]
    keep = np.abs(k - centre[..., None]) < w[..., None]
    keep &= g.random((R0, R1, 1)) > 0.2                      # some columns have no points at all
    return np.flatnonzero(keep.reshape(-1)).astype(np.int64)


RUN_CASES = [
    # lattice, ids, chunk tiles, calib            what it exercises
    ((8, 8, 64), "all", None, "default"),         # short columns: two runs per tile, every run whole
    ((16, 16, 96), "band12", None, "default"),    # runs of 1..24 rows, ragged last tile
    ((16, 16, 96), "band12", None, "scaled"),     # some columns out of the fine bounding box (masked to 0)
    ((24, 24, 128), "band20", 8, "default"),      # chunks of 1024 rows: runs cut at chunk boundaries, many launches
    ((48, 48, 8), "random", None, "default"),     # ~1.3 rows per column: nearly one segment per row
    ((5, 3, 64), "band3", None, "default"),       # fewer than 128 rows: one partial tile
]


@pytest.mark.parametrize("R,kind,tiles,calib_name", RUN_CASES, ids=[c[1] + "-" + "x".join(map(str, c[0])) + ("-" + c[3] if c[3] != "default" else "") for c in RUN_CASES])
def test_chain_runlist_vs_oracle(setup, R, kind, tiles, calib_name):
    prob, fine, netMR, eng = setup
    calib = syn.default_calib() if calib_name == "default" else syn.scaled_calib()
    netMR.query(syn.random_points(256).cuda(), calib.cuda())     # the engine is per device: re-snapshot this net's MLPs
    total = R[0] * R[1] * R[2]
    if kind == "all":
        ids = np.arange(total, dtype=np.int64)
    elif kind == "random":
        ids = np.sort(np.random.default_rng(5).choice(total, 3000, replace=False)).astype(np.int64)
    else:
        ids = band_ids(R, int(kind[4:]), 11)
    ref = orc.query_fine(fine, lattice_points(R, calib, ids), calib)[0].numpy().ravel()
    tid = torch.from_numpy(ids)
    if tiles:
        eng.set_chunk_tiles(tiles)
    try:
        eng.set_chain(1)
        eng.profile(True)
        out = eng.eval_lattice_ids(2, R, tid, calib[0]).cpu().numpy()
        n_rows_launches = eng.profile_read_kind(2)[0]
        eng.profile(False)
        eng.set_chain(2)
        layer = eng.eval_lattice_ids(2, R, tid, calib[0]).cpu().numpy()
    finally:
        eng.set_chain(1)
        if tiles:
            eng.set_chunk_tiles(16 * 148)
    # every id list takes the run-list chain whatever its run lengths (values must not depend on how a
    # caller cuts a list into calls); launches are whole 1024-row blocks, at most `tiles` tiles each
    assert n_rows_launches == (-(-len(ids) // (tiles * 128)) if tiles else 1), "run-list chain not taken"
    assert np.abs(out - ref).max() < OCC_TOL
    assert np.array_equal(out == 0, ref == 0)              # identical in-bounds masks
    assert np.abs(out - layer).max() < OCC_TOL


def test_chain_runlist_unsorted_ids(setup):
    """Runs are found by adjacency only: an unsorted list is still evaluated point by point."""
    prob, fine, netMR, eng = setup
    R = (16, 16, 96)
    calib = syn.default_calib()
    netMR.query(syn.random_points(256).cuda(), calib.cuda())
    ids = band_ids(R, 12, 3)
    g = np.random.default_rng(0)
    blocks = np.array_split(ids, 40)
    g.shuffle(blocks)
    ids = np.concatenate(blocks)
    ref = orc.query_fine(fine, lattice_points(R, calib, ids), calib)[0].numpy().ravel()
    out = eng.eval_lattice_ids(2, R, torch.from_numpy(ids), calib[0]).cpu().numpy()
    assert np.abs(out - ref).max() < OCC_TOL
