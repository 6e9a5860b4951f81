"""CPU: host-side logic of the drop-in mesh_util (callback forms) against the oracle restatement
and the reference-generated golden checksums."""
import hashlib
import io
import os

import numpy as np
import pytest
import torch

from helpers import golden, orc
from test_oracle_golden import ANALYTIC

from pifu_b200 import mesh_util


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_create_grid_matches_oracle_lattice():
    coords, mat = mesh_util.create_grid(16, 16, 16)
    oc, omat, _ = orc.lattice_coords(16, torch.eye(4)[None])
    assert np.array_equal(coords, oc) and np.array_equal(mat, omat)


@pytest.mark.parametrize("name", ["ellipsoid", "ripple"])
@pytest.mark.parametrize("res,init", [(64, 8), (96, 12), (128, 32)])
def test_host_octree_bit_exact_vs_reference(name, res, init):
    g = golden("octree_analytic.npz")
    coords, _ = mesh_util.create_grid(res, res, res)
    calls = []
    f = mesh_util.eval_grid_octree(coords, lambda p: (calls.append(p.shape[1]), ANALYTIC[name](p))[1],
                                   init_resolution=init, num_samples=50000)
    key = "%s_%d_%d" % (name, res, init)
    assert sum(calls) == int(g[key + "_evaluated"])
    assert np.array_equal(sha(f), g[key + "_sha"])


def test_fill_gather_equals_sequential_loop():
    """The per-voxel restatement used by the CUDA fill kernel == the reference's loop order."""
    rng = np.random.default_rng(3)
    R, step = 24, 4
    nc = R // step - 1
    for _ in range(5):
        skip = rng.uniform(size=(nc,) * 3) < 0.3
        mid = rng.uniform(size=(nc,) * 3)
        sdf_a = rng.uniform(size=(R,) * 3)
        todo_a = rng.uniform(size=(R,) * 3) < 0.5
        sdf_b, todo_b = sdf_a.copy(), todo_a.copy()
        for cx, cy, cz in zip(*np.where(skip)):
            x, y, z = cx * step, cy * step, cz * step
            sdf_a[x:x + step + 1, y:y + step + 1, z:z + step + 1] = mid[cx, cy, cz]
            todo_a[x:x + step + 1, y:y + step + 1, z:z + step + 1] = False
        mesh_util._fill_from_skip_cells(sdf_b, todo_b, skip, mid, step)
        assert np.array_equal(sdf_a, sdf_b) and np.array_equal(todo_a, todo_b)
        sdf_c, todo_c = orc.octree_fill_gather(sdf_a.copy(), todo_a.copy(), skip, mid, step)
        assert np.array_equal(sdf_c, sdf_a)


def test_save_obj_format(tmp_path):
    v = np.array([[0.123456, -1.5, 2.0], [1, 2, 3]], dtype=np.float64)
    c = np.array([[0.5, 0.25, 1.0], [0, 0, 0]])
    f = np.array([[0, 1, 1]], dtype=np.int32)
    p = str(tmp_path / "m.obj")
    mesh_util.save_obj_mesh_with_color(p, v, f, c)
    lines = open(p).read().splitlines()
    assert lines[0] == "v 0.1235 -1.5000 2.0000 0.5000 0.2500 1.0000"
    assert lines[2] == "f 1 2 2"
