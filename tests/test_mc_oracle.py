"""CPU: properties of the marching-cubes oracle (the table generator's watertightness claim,
orientation, interpolation) - the oracle is what the CUDA kernel is held to bit for bit."""
import numpy as np
import pytest

from oracle import mc_oracle


def sphere(n, r=0.6, c=(0.03, -0.02, 0.05)):
    g = np.linspace(-1, 1, n)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    return (r - np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2)).astype(np.float32) + 0.5


def edges_of(faces):
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)
    return e


def test_sphere_closed_oriented_manifold():
    vol = sphere(40)
    v, f, n, val, cases = mc_oracle.marching_cubes(vol, 0.5)
    e = edges_of(f)
    # every directed edge appears once and its reverse once: closed, consistently oriented
    key = e[:, 0].astype(np.int64) * len(v) + e[:, 1]
    rkey = e[:, 1].astype(np.int64) * len(v) + e[:, 0]
    assert len(np.unique(key)) == len(key)
    assert np.array_equal(np.sort(key), np.sort(rkey))
    # Euler characteristic of a sphere
    assert len(v) - len(key) // 2 + len(f) == 2
    # normals of the faces point towards decreasing values (outwards for this field)
    c = (np.array(vol.shape) - 1) / 2 + np.array([0.03, -0.02, 0.05]) * (vol.shape[0] - 1) / 2
    tri = v[f]
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    out = tri.mean(1) - c
    assert ((fn * out).sum(1) > 0).mean() > 0.999
    assert ((n * (v - c)).sum(1) > 0).all()
    # vertices sit on the iso-surface to interpolation accuracy
    r = np.linalg.norm((v - c) * 2 / (vol.shape[0] - 1), axis=1)
    assert np.abs(r - 0.6).max() < 2e-3


def test_random_field_watertight():
    """Noise field: every ambiguous configuration occurs; interior edges must still pair up."""
    rng = np.random.default_rng(0)
    vol = rng.uniform(0, 1, (14, 15, 16)).astype(np.float32)
    vol[0] = vol[-1] = 0
    vol[:, 0] = vol[:, -1] = 0
    vol[:, :, 0] = vol[:, :, -1] = 0          # surface cannot leave the volume -> closed
    v, f, _, _, cases = mc_oracle.marching_cubes(vol, 0.5)
    assert len(np.unique(cases)) > 200
    e = edges_of(f)
    key = e[:, 0].astype(np.int64) * len(v) + e[:, 1]
    rkey = e[:, 1].astype(np.int64) * len(v) + e[:, 0]
    assert len(np.unique(key)) == len(key)
    assert np.array_equal(np.sort(key), np.sort(rkey))


def test_first_use_numbering_and_errors():
    vol = sphere(12)
    v, f, _, _, _ = mc_oracle.marching_cubes(vol, 0.5)
    # vertex ids appear in increasing order of first use in the face list
    first = {}
    for idx in f.reshape(-1):
        first.setdefault(int(idx), len(first))
    assert all(k == i for k, i in first.items())
    with pytest.raises(ValueError):
        mc_oracle.marching_cubes(vol, 5.0)
    flat = np.full((6, 6, 6), 0.25, np.float32)
    with pytest.raises((ValueError, RuntimeError)):
        mc_oracle.marching_cubes(flat, 0.5)


# ----------------------------------------------------------------------------- the rule, the tables, the oracle
def _generator():
    import importlib.util
    import os
    from helpers import ROOT
    path = os.path.join(ROOT, "rgb-d-pifuhd_b200", "tools", "gen_mc_tables.py")
    spec = importlib.util.spec_from_file_location("gen_mc_tables", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _header_rows():
    """(MC_AMB, MC_SUB_BASE, MC_TRIS rows) parsed from the header the CUDA kernels compile."""
    import os
    import re
    from helpers import ROOT
    text = open(os.path.join(ROOT, "rgb-d-pifuhd_b200", "csrc", "mc_tables.h")).read()

    def flat(name):
        m = re.search(name + r"\[[^\]]*\](?:\[[^\]]*\])? = \{(.*?)\};", text, re.S)
        return [int(x) for x in re.findall(r"-?\d+", m.group(1))]
    nsub = int(re.search(r"#define MC_NSUB (\d+)", text).group(1))
    max_t = int(re.search(r"#define MC_MAX_TRIS (\d+)", text).group(1))
    tris = np.array(flat("MC_TRIS")).reshape(nsub, 3 * max_t)
    return flat("MC_AMB"), flat("MC_SUB_BASE"), [[int(x) for x in r if x >= 0] for r in tris]


def test_header_is_the_generators_output():
    g = _generator()
    amb, base, rows = g.build()
    h_amb, h_base, h_rows = _header_rows()
    assert h_amb == amb and h_base == base
    assert h_rows == [t for t, _ in rows]


def test_oracle_derives_the_table_rows_independently():
    """oracle/mc_ref.c includes no table: it applies the rule to a cell's eight values at run time.  For random cells
    (all 254 surface cases, every reachable resolution of their ambiguous faces) its triangles must be the row the
    CUDA kernels would read - so a wrong table entry, sub-case index or face test cannot hide."""
    g = _generator()
    amb, base, rows = _header_rows()
    rng = np.random.default_rng(0)
    seen = set()
    for it in range(60000):
        case = int(rng.integers(1, 255)) if it >= 254 else it + 1
        mag = np.exp(rng.uniform(-3, 3, 8))
        inside = np.array([(case >> c) & 1 for c in range(8)], bool)
        v = np.where(inside, 0.5 + mag, 0.5 - mag)
        bits = q = 0
        for fi, f in enumerate(g.FACES):
            if (amb[case] >> fi) & 1:
                i0 = 0 if inside[f[0]] else 1
                pin = (v[f[i0]] - 0.5) * (v[f[i0 + 2]] - 0.5)
                pout = (v[f[i0 ^ 1]] - 0.5) * (v[f[(i0 ^ 1) + 2]] - 0.5)
                bits |= int(pin > pout) << q
                q += 1
        sub = base[case] + bits
        assert mc_oracle.cell_triangles(v, 0.5) == rows[sub], (case, bits)
        seen.add(sub)
    assert len(seen) > 600                                  # of 656 rows; the rest need value patterns that cannot occur


def test_table_properties_the_kernels_pack_on():
    """csrc/mc.cu packs a surface cell into one 32-bit word (table row 10 bits, triangles 4, owned vertices 4, up to three
    owned edges 4 bits each) and the row pass writes an interior cell's vertex jobs from that word alone.  That is sound
    only if, for every table row: the row index fits 10 bits, the triangle and vertex counts fit 4, and a cell none of
    whose low faces lies on the volume border is the first user of at most the three edges at its high corner (5, 6, 10:
    the edges whose LOWMASK is 0), which all appear in the row's vertex list when the case cuts them."""
    import os
    import re
    from helpers import ROOT
    text = open(os.path.join(ROOT, "rgb-d-pifuhd_b200", "csrc", "mc_tables.h")).read()

    def flat(name):
        m = re.search(name + r"\[[^\]]*\](?:\[[^\]]*\])? = \{(.*?)\};", text, re.S)
        return [int(x) for x in re.findall(r"-?\d+", m.group(1))]
    nsub = int(re.search(r"#define MC_NSUB (\d+)", text).group(1))
    max_t = int(re.search(r"#define MC_MAX_TRIS (\d+)", text).group(1))
    assert nsub <= 1024 and max_t <= 15
    ntri, nvert, low = flat("MC_NTRI"), flat("MC_NVERT"), flat("MC_EDGE_LOWMASK")
    verts = np.array(flat("MC_VERTS")).reshape(nsub, -1)
    sub_case = flat("MC_SUB_CASE")
    assert len(ntri) == nsub and len(nvert) == nsub and len(low) == 12 and len(sub_case) == nsub
    assert [e for e in range(12) if low[e] == 0] == [5, 6, 10]
    g = _generator()
    for r in range(nsub):
        if sub_case[r] in (0, 255):                             # no surface: never packed
            assert ntri[r] == 0 and nvert[r] == 0
            continue
        assert 1 <= ntri[r] <= max_t and 3 <= nvert[r] <= 12    # (a surface cell's word is never 0: the row pass relies on it)
        vs = [int(e) for e in verts[r][:nvert[r]]]
        assert len(set(vs)) == len(vs) and all(0 <= e < 12 for e in vs)
        # the row's vertices are exactly the edges its case cuts (so the owned COUNT depends on the case only)
        case = sub_case[r]
        cut = sorted(e for e in range(12) if ((case >> g.EDGES[e][0]) & 1) != ((case >> g.EDGES[e][1]) & 1))
        assert sorted(vs) == cut, (r, case)
        for zm in range(8):
            owned = [e for e in vs if (low[e] & ~zm) == 0]
            assert len(owned) <= 12
            if zm == 0:
                assert len(owned) <= 3 and set(owned) <= {5, 6, 10}


def test_face_test_is_the_asymptotic_decider():
    """One ambiguous face, bilinear values: the inside corners are joined across the face exactly when the
    interpolant's saddle value is inside (Lewiner's test_face / Nielson-Hamann)."""
    g = _generator()
    rng = np.random.default_rng(3)
    case = (1 << 0) | (1 << 2)                              # corners 0 and 2: a diagonal of face 0, also cut on other faces
    for _ in range(500):
        a, c = 0.5 + rng.uniform(0.01, 1, 2)
        b, d = 0.5 - rng.uniform(0.01, 1, 2)
        v = np.array([a, b, c, d, 0.1, 0.1, 0.1, 0.1])
        tris = mc_oracle.cell_triangles(v, 0.5)
        saddle = ((a - 0.5) * (c - 0.5) - (b - 0.5) * (d - 0.5)) / ((a - 0.5) + (c - 0.5) - (b - 0.5) - (d - 0.5))
        edges_used = set(tris)
        # separated: two triangles (one per inside corner); joined: the corners share one sheet (4 triangles, 6 vertices)
        assert (len(tris) // 3 == 4) == (saddle > 0), (v, tris)
        assert edges_used == {0, 3, 8, 1, 2, 10}


def test_torus_euler_characteristic():
    n = 48
    gx = np.linspace(-1.3, 1.3, n)
    x, y, z = np.meshgrid(gx, gx, gx, indexing="ij")
    vol = (0.25 - np.sqrt((np.sqrt(x ** 2 + y ** 2) - 0.8) ** 2 + z ** 2)).astype(np.float32) + 0.5
    v, f, _, _, _ = mc_oracle.marching_cubes(vol, 0.5)
    e = edges_of(f)
    key = e[:, 0].astype(np.int64) * len(v) + e[:, 1]
    rkey = e[:, 1].astype(np.int64) * len(v) + e[:, 0]
    assert np.array_equal(np.sort(key), np.sort(rkey)) and len(np.unique(key)) == len(key)
    assert len(v) - len(key) // 2 + len(f) == 0              # genus 1


@pytest.mark.parametrize("name", ["random14", "blob24", "ragged_9_12_17"])
def test_oracle_reproduces_committed_fixture(name):
    """tests/golden/mc_rule.npz (oracle/make_golden_mc.py) pins the rule: the oracle built here must give the committed
    meshes bit for bit (float64 positions included: -ffp-contract=off).  The GPU suite compares the kernels with the same
    file.  (Not a scikit-image fixture - that package is absent; see the header of the generating script.)"""
    import hashlib
    from helpers import golden
    g = golden("mc_rule.npz")
    v, f, n, val = mc_oracle.marching_cubes(g[name + "_volume"], float(g[name + "_level"]))[:4]
    assert np.array_equal(f, g[name + "_faces"])
    assert np.array_equal(v, g[name + "_verts"])
    assert hashlib.sha256(np.ascontiguousarray(n).tobytes()).hexdigest() == str(g[name + "_normals_sha256"])
    assert hashlib.sha256(np.ascontiguousarray(val).tobytes()).hexdigest() == str(g[name + "_values_sha256"])
