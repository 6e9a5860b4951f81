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
