"""GPU parity of the device octree (bit-exact field vs the reference loop) and of the CUDA
marching cubes (bit-exact topology and vertices vs the sequential CPU oracle)."""
import hashlib

import numpy as np
import pytest
import torch

from helpers import calibrated_problem, golden, oracle_states, orc, syn
from test_oracle_golden import ANALYTIC
from test_query_gpu import build_nets, sign_agreement

pytestmark = pytest.mark.gpu


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.fixture(scope="module")
def sat():
    torch.set_grad_enabled(False)
    prob, _ = calibrated_problem(saturated=True)
    netG, netMR = build_nets(prob)
    return prob, netG, netMR


def _analytic_on_device(name, ids, res):
    """The golden generator's analytic fields evaluated on device lattice ids in float64 with
    IEEE-exact ops only (+,-,*,/,sqrt,clip) -> bit-identical to the numpy evaluation."""
    k = (ids % res).double()
    j = ((ids // res) % res).double()
    i = (ids // (res * res)).double()
    step = 2.0 / res
    x, y, z = i * step + (-1.0), j * step + (-1.0), k * step + (-1.0)
    if name == "ellipsoid":
        r = torch.sqrt((x / 0.35) ** 2 + (y / 0.8) ** 2 + (z / 0.3) ** 2)
        v = 0.5 + 2.0 * (1.0 - r)
    else:
        v = 0.5 + 0.6 * (x * y - z * z) + 0.25 * (x * x * x - y * z)
    return torch.clamp(v, 0.0, 1.0).float()


@pytest.mark.parametrize("name", ["ellipsoid", "ripple"])
@pytest.mark.parametrize("res,init", [(64, 8), (64, 16), (96, 12), (128, 32), (128, 64)])
def test_device_octree_bit_exact_vs_reference_loop(name, res, init):
    """Frontier compaction + skip + fill on the device reproduce the reference's float64 field
    bit for bit (sha256 recorded from mesh_util.eval_grid_octree itself) and evaluate exactly
    the same number of lattice points."""
    from pifu_b200 import get_engine
    g = golden("octree_analytic.npz")
    eng = get_engine("cuda")
    eng.octree_begin(res, init, 0.05)
    evaluated = 0
    while True:
        step, ids = eng.octree_frontier()
        if step == 0:
            break
        assert bool((ids[1:] > ids[:-1]).all())          # C order of the boolean mask
        evaluated += ids.numel()
        eng.octree_commit(_analytic_on_device(name, ids, res))
    sdf64, sdf32 = eng.octree_export(want64=True, want32=True)
    key = "%s_%d_%d" % (name, res, init)
    f = sdf64.cpu().numpy()
    # numpy's x**2 and torch's x**2 are both exact products; cross-check a probe before the hash
    assert np.array_equal(f[::7, ::5, ::3], g[key + "_probe"])
    assert evaluated == int(g[key + "_evaluated"])
    assert np.array_equal(sha(f), g[key + "_sha"])
    assert np.array_equal(sdf32.cpu().numpy(), f.astype(np.float32))


@pytest.mark.parametrize("name", ["ellipsoid", "ripple"])
@pytest.mark.parametrize("res,init", [(64, 8), (96, 12), (128, 32), (32, 64)])
def test_callback_octree_bit_exact_vs_reference(name, res, init):
    """mesh_util.eval_grid_octree(coords, eval_func) with an arbitrary host callable: device
    bookkeeping, float64 values committed as they are -> the reference's field bit for bit
    (sha256 recorded from the reference's own loop) and the same chunked calls."""
    from pifu_b200 import mesh_util
    coords, _ = mesh_util.create_grid(res, res, res)
    calls = []
    f = mesh_util.eval_grid_octree(coords, lambda p: (calls.append(p.shape[1]), ANALYTIC[name](p))[1],
                                   init_resolution=init, num_samples=50000)
    assert f.dtype == np.float64 and f.shape == (res, res, res)
    if init > res:                      # `mesh_util.py:138`: reso = 0, the loop never runs
        assert not f.any() and not calls
        return
    g = golden("octree_analytic.npz")
    key = "%s_%d_%d" % (name, res, init)
    assert sum(calls) == int(g[key + "_evaluated"]) and max(calls) <= 50000
    assert np.array_equal(sha(f), g[key + "_sha"])
    # a float64-valued callable keeps its precision in the field (reference: sdf[test_mask] = values)
    f64 = mesh_util.eval_grid_octree(coords, lambda p: 0.5 + 0.1 * p[0] + 1e-12 * p[1], init_resolution=init)
    ref = orc.eval_grid_octree(coords, lambda p: 0.5 + 0.1 * p[0] + 1e-12 * p[1], init_resolution=init)
    assert np.array_equal(f64, ref)


def test_octree_with_net(sat):
    """Same evaluated values -> identical field: run the oracle's restatement of the reference
    loop on the GPU's own dense field and compare with the device octree bit for bit; then
    compare with the reference's field (golden) within the 16-bit-operand tolerance."""
    prob, netG, netMR = sat
    eng = netMR._engine_for(torch.zeros(1, device="cuda"))
    eng.sync_features(0, netG.im_feat_list[-1])
    eng.sync_features(1, netMR.im_feat_list[-1])
    calib = syn.default_calib()
    res = 64
    # every lattice point through the path the octree's frontiers take (sorted id list -> run-list chain
    # kernel): a point's value does not depend on which other points share its call
    dense = eng.eval_lattice_ids(2, res, torch.arange(res ** 3), calib).cpu().numpy().astype(np.float32)
    assert np.abs(dense - eng.eval_grid(2, res, calib).cpu().numpy()).max() < 8e-3
    coords, _, _ = orc.lattice_coords(res, calib)

    def lookup(points):
        idx = np.rint((points + 1.0) * (res / 2.0)).astype(np.int64)
        idx[1] = np.rint((-points[1] + 1.0) * (res / 2.0)).astype(np.int64)     # calib flips y
        return dense[(idx[0] * res + idx[1]) * res + idx[2]]
    stats = []
    ref = orc.eval_grid_octree(coords, lookup, init_resolution=16, num_samples=10 ** 9, stats=stats)
    sdf64, _, ev = eng.eval_grid_octree(2, res, calib, init_resolution=16, want64=True)
    assert ev == [n for _, n in stats]
    assert np.array_equal(sdf64.cpu().numpy(), ref)
    gold = golden("query_none.npz")["mr_octree64_init16"]
    out = sdf64.cpu().numpy()
    assert np.abs(out - gold).max() < 0.03          # filled midpoints inherit two corner errors
    assert sign_agreement(out, gold) >= 0.9995
    assert sum(ev) < 0.6 * res ** 3                  # the octree actually prunes on this field


def _oracle_mc(vol, level=0.5):
    from oracle import mc_oracle
    return mc_oracle.marching_cubes(vol, level)


def _sphere(n):
    g = np.linspace(-1, 1, n)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    return (0.6 - np.sqrt((x - 0.03) ** 2 + (y + 0.02) ** 2 + (z - 0.05) ** 2)).astype(np.float32) + 0.5


def _noise(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(0, 1, shape).astype(np.float32)


@pytest.mark.parametrize("vol", ["sphere33", "sphere64", "noise", "noise_ragged", "tiny", "noise_wide", "noise_aligned"])
def test_marching_cubes_bit_exact(vol):
    from pifu_b200 import get_engine
    # noise_wide: rows of 1 099 cells = 138 threads' worth -> the row pass runs teams of 64 threads on named barriers and
    # the classify pass several 128-value tiles per row; noise_aligned: 16-byte aligned rows (the vector-load paths)
    v = {"sphere33": _sphere(33), "sphere64": _sphere(64), "noise": _noise((24, 24, 24), 1),
         "noise_ragged": _noise((9, 17, 31), 2), "tiny": _noise((2, 2, 2), 5), "noise_wide": _noise((4, 5, 1100), 7),
         "noise_aligned": _noise((19, 12, 256), 8)}[vol]
    if vol == "tiny":
        v[0, 0, 0], v[1, 1, 1] = 0.9, 0.1
    rv, rf, rn, rval, _ = _oracle_mc(v)
    eng = get_engine("cuda")
    verts, faces, normals, values = eng.marching_cubes(torch.from_numpy(v).cuda(), 0.5)
    assert np.array_equal(faces.cpu().numpy(), rf)                    # topology + numbering: bit-exact
    assert np.array_equal(verts.cpu().numpy(), rv)                    # float64 positions: bit-exact
    assert np.abs(verts.cpu().numpy() - rv).max() <= 1e-5             # (north_star's stated tolerance)
    assert np.array_equal(values.cpu().numpy(), rval)
    assert np.abs(normals.cpu().numpy() - rn).max() < 1e-6


@pytest.mark.parametrize("name", ["random14", "blob24", "ragged_9_12_17"])
def test_marching_cubes_committed_fixture(name):
    """The kernels against the COMMITTED meshes of tests/golden/mc_rule.npz (oracle/make_golden_mc.py; pinned on CPU by
    tests/test_mc_oracle.py::test_oracle_reproduces_committed_fixture) - no oracle code runs here."""
    from pifu_b200 import get_engine
    g = golden("mc_rule.npz")
    vol, level = g[name + "_volume"], float(g[name + "_level"])
    verts, faces, normals, values = get_engine("cuda").marching_cubes(torch.from_numpy(vol).cuda(), level)
    assert np.array_equal(faces.cpu().numpy(), g[name + "_faces"])
    assert np.array_equal(verts.cpu().numpy(), g[name + "_verts"])
    assert hashlib.sha256(values.cpu().numpy().tobytes()).hexdigest() == str(g[name + "_values_sha256"])


def test_marching_cubes_no_surface():
    from pifu_b200 import get_engine
    eng = get_engine("cuda")
    with pytest.raises(ValueError):
        eng.marching_cubes(torch.full((8, 8, 8), 0.25, device="cuda"), 0.5)


def test_reconstruction_end_to_end(sat):
    """mesh_util.reconstruction (reference signature): device octree + device MC; the mesh must
    equal the CPU oracle's MC of the very same field, transformed as mesh_util.py:87-92."""
    from pifu_b200 import mesh_util
    prob, netG, netMR = sat
    calib = syn.default_calib().cuda()
    res = 64
    out = mesh_util.reconstruction(netMR, "cuda", calib, res, None, None, thresh=0.5, use_octree=True,
                                   num_samples=5000)
    assert out != -1
    verts, faces, normals, values = out
    assert verts.dtype == np.float64 and faces.dtype == np.int32 and normals.dtype == np.float32
    eng = netMR._engine_for(torch.zeros(1, device="cuda"))
    _, sdf32, _ = eng.eval_grid_octree(2, res, calib[0], want64=False, want32=True)
    rv, rf, _, _, _ = _oracle_mc(sdf32.cpu().numpy())
    mat = np.eye(4)
    mat[0, 0] = mat[1, 1] = mat[2, 2] = 2.0 / res
    mat[:3, 3] = -1
    trans = np.linalg.inv(calib[0].cpu().numpy()) @ mat
    rv = (trans[:3, :3] @ rv.T + trans[:3, 3:4]).T
    assert np.linalg.det(trans[:3, :3]) < 0
    assert np.array_equal(faces, rf[:, ::-1])
    assert np.abs(verts - rv).max() < 1e-12
    # dense evaluation gives the same surface wherever the octree did not skip
    dense = mesh_util.reconstruction(netMR, "cuda", calib, res, None, None, use_octree=False)
    assert dense != -1 and abs(len(dense[0]) - len(verts)) < 0.2 * len(verts)
    # an empty field follows the reference's error convention
    flat = mesh_util.reconstruction(_Flat(), "cuda", calib, 16, None, None, use_octree=False)
    assert flat == -1


class _Flat:
    """A foreign `net` (not this package's): generic callback path of reconstruction()."""

    def query(self, samples, calib):
        self.preds = torch.full((1, 1, samples.shape[2]), 0.25, device=samples.device)

    def get_preds(self):
        return self.preds


@pytest.mark.parametrize("vol,splits", [("sphere33", [0, 11, 22, 33]), ("noise", [0, 6, 7, 20, 24]),
                                        ("noise_ragged", [0, 2, 4, 9]), ("sphere64", [0, 32, 64])])
def test_marching_cubes_slabs_assemble_to_whole(vol, splits):
    """Slab form (multi-GPU sharding along axis 0, run here slab after slab on one GPU): each slab
    bit-exact against the oracle's slab form, and the assembled fragments bit-exact against the
    oracle's traversal of the whole volume."""
    from oracle import mc_oracle
    from pifu_b200 import get_engine
    v = {"sphere33": _sphere(33), "sphere64": _sphere(64), "noise": _noise((24, 24, 24), 1),
         "noise_ragged": _noise((9, 17, 31), 2)}[vol]
    R0 = v.shape[0]
    rv, rf, rn, rval, _ = _oracle_mc(v)
    eng = get_engine("cuda")
    dv = torch.from_numpy(v).cuda()
    V, F, N, VAL, first = [], [], [], [], 0
    for pb, pe in zip(splits[:-1], splits[1:]):
        ce, lo, hi = min(pe, R0 - 1), max(pb - 1, 0), min(pe + 2, R0)
        if ce <= pb:
            continue
        sv, sf, sn, sval, ng = eng.marching_cubes_slab(dv[lo:hi], 0.5, lo, R0, ce - lo, pb > 0)
        ov, of, on, oval, ong = mc_oracle.marching_cubes_slab(v[lo:hi], 0.5, lo, R0, ce - lo, pb > 0)
        # (the first ng vertices are the ghost layer's: numbered, owned and computed by the slab before, not written here)
        assert ng == ong and np.array_equal(sf.cpu().numpy(), of) and np.array_equal(sv.cpu().numpy()[ng:], ov[ng:])
        assert np.array_equal(sval.cpu().numpy()[ng:], oval[ng:]) and np.abs(sn.cpu().numpy()[ng:] - on[ng:]).max() < 1e-6
        V.append(sv[ng:]); N.append(sn[ng:]); VAL.append(sval[ng:]); F.append(sf + (first - ng))
        first += sv.shape[0] - ng
    assert np.array_equal(torch.cat(F).cpu().numpy(), rf)
    assert np.array_equal(torch.cat(V).cpu().numpy(), rv)
    assert np.array_equal(torch.cat(VAL).cpu().numpy(), rval)
    assert np.abs(torch.cat(N).cpu().numpy() - rn).max() < 1e-6


def test_marching_cubes_extract_overflow_retries():
    """`pifu_mc_extract` sizes nothing on the host: outputs are allocated from a hint and only the counts come back.
    A hint that is too small must leave the outputs untouched and the retry must give the same mesh."""
    from pifu_b200 import get_engine
    eng = get_engine("cuda")
    v = _noise((20, 21, 22), 9)
    rv, rf, _, _, _ = _oracle_mc(v)
    key = ("whole",) + v.shape
    eng._mc_hint[key] = (1, 1)                    # capacity ~1 k vertices / ~2 k faces: far too small
    verts, faces, _, _ = eng.marching_cubes(torch.from_numpy(v).cuda(), 0.5)
    assert len(rv) > 2048 and np.array_equal(faces.cpu().numpy(), rf) and np.array_equal(verts.cpu().numpy(), rv)
    assert eng._mc_hint[key] == (len(rv), len(rf))
    verts2, faces2, _, _ = eng.marching_cubes(torch.from_numpy(v).cuda(), 0.5)      # sized from the hint: one pass
    assert np.array_equal(faces2.cpu().numpy(), rf) and np.array_equal(verts2.cpu().numpy(), rv)
