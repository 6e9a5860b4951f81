"""CPU ORACLE (test infrastructure): writes tests/golden/mc_rule.npz - the marching-cubes oracle's own output on three
small seeded fields.  NOT a reference-generated fixture: scikit-image is absent here ("parity unpinned"),
so this pins the *rule* (rgb-d-pifuhd_b200/tools/gen_mc_tables.py) against silent changes of the oracle, the generator or the
compiler's floating point, nothing more.  Run from the repository root:  python oracle/make_golden_mc.py"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mc_oracle                      # noqa: E402


def fields():
    rng = np.random.default_rng(20260117)
    yield "random14", rng.random((14, 14, 14)).astype(np.float32), 0.5              # every case, every ambiguous face
    g = np.linspace(-1, 1, 24)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    yield "blob24", (0.7 + 0.1 * np.sin(5 * x) * np.cos(4 * y) - np.sqrt(x * x + 1.3 * y * y + 0.8 * z * z)).astype(np.float32), 0.0
    yield "ragged_9_12_17", rng.normal(size=(9, 12, 17)).astype(np.float32), 0.25   # unequal, non-multiple-of-4 extents


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {}
    for name, vol, level in fields():
        v, f, n, val = mc_oracle.marching_cubes(vol, level)[:4]
        out[name + "_volume"] = vol
        out[name + "_level"] = np.float64(level)
        out[name + "_verts"] = v
        out[name + "_faces"] = f
        out[name + "_normals_sha256"] = np.array(digest(n))
        out[name + "_values_sha256"] = np.array(digest(val))
        print(name, v.shape, f.shape)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mc_rule.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
