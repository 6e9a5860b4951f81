"""The order-free restatement of the reference's sequential octree fill (`mesh_util.py:181-184`) that
csrc/octree.cu implements, checked on the CPU against the loop itself on random skip masks.

Reference: for every skip cell in C order, `sdf[x:x+s+1, y:y+s+1, z:z+s+1] = mid` (inclusive range, clipped by the
array; later cells overwrite earlier ones).  Restatements:
  * voxel-centric (`fill_kernel`): a voxel takes the value of the lexicographically largest skip cell covering it;
    per axis the candidates are p // s and, when p % s == 0, p // s - 1;
  * cell-centric (`fill_cells_kernel`): a skip cell writes every voxel of its block except those that a skip cell
    visited LATER (first non-zero offset +1 among the 26 neighbours) also covers; a neighbour at offset o covers the
    voxel at local offset d iff for every axis o_a == 0, or o_a == +1 and d_a == s, or o_a == -1 and d_a == 0."""
import itertools

import numpy as np
import pytest


def sequential(skip, mid, s, R):
    out = np.full((R, R, R), np.nan)
    for x, y, z in zip(*np.nonzero(skip)):              # np.nonzero is C order, like the reference's loop over skip cells
        out[x * s:x * s + s + 1, y * s:y * s + s + 1, z * s:z * s + s + 1] = mid[x, y, z]
    return out


def voxel_centric(skip, mid, s, R):
    out = np.full((R, R, R), np.nan)
    c = skip.shape[0]
    for p in itertools.product(range(R), repeat=3):
        cands = []
        for a in range(3):
            hi = p[a] // s
            ca = [hi] if hi < c else []
            if p[a] % s == 0 and hi - 1 >= 0 and hi - 1 < c:
                ca.append(hi - 1)
            cands.append(ca)                            # descending per axis
        for cell in itertools.product(*cands):          # descending lexicographic order
            if skip[cell]:
                out[p] = mid[cell]
                break
    return out


def cell_centric(skip, mid, s, R):
    out = np.full((R, R, R), np.nan)
    c = skip.shape[0]
    writes = np.zeros((R, R, R), dtype=int)
    offsets = [o for o in itertools.product((-1, 0, 1), repeat=3)
               if next((v for v in o if v != 0), 0) > 0]                      # visited later by the C-order loop
    for x, y, z in zip(*np.nonzero(skip)):
        later = [o for o in offsets
                 if all(0 <= q < c for q in (x + o[0], y + o[1], z + o[2])) and skip[x + o[0], y + o[1], z + o[2]]]
        for d in itertools.product(range(s + 1), repeat=3):
            p = (x * s + d[0], y * s + d[1], z * s + d[2])
            if max(p) >= R:
                continue
            covered_later = any(all(o[a] == 0 or (o[a] == 1 and d[a] == s) or (o[a] == -1 and d[a] == 0) for a in range(3))
                                for o in later)
            if not covered_later:
                out[p] = mid[x, y, z]
                writes[p] += 1
    assert writes.max() <= 1, "two skip cells claim the same voxel"        # the CUDA kernel relies on this (no write race)
    return out


@pytest.mark.parametrize("s,cells,density,seed", [(2, 5, 0.5, 0), (2, 6, 0.85, 1), (4, 3, 0.6, 2), (3, 4, 0.4, 3), (2, 4, 1.0, 4)])
def test_fill_restatements_equal_the_reference_loop(s, cells, density, seed):
    rng = np.random.default_rng(seed)
    R = cells * s + (1 if seed % 2 else 2)              # R - 1 a multiple of the stride or not: the clipped last range
    c = -(-R // s) - 1                                  # cells per axis, `mesh_util.py:154-160`
    skip = rng.random((c, c, c)) < density
    mid = rng.random((c, c, c))
    ref = sequential(skip, mid, s, R)
    np.testing.assert_array_equal(voxel_centric(skip, mid, s, R), ref)
    np.testing.assert_array_equal(cell_centric(skip, mid, s, R), ref)
