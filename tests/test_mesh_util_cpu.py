"""CPU: host-side logic of the drop-in mesh_util (lattice, callback loop, OBJ writer) against the
oracle restatement; the octree forms need the device and are covered by the GPU tests."""
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
        orc.octree_fill_gather(sdf_b, todo_b, skip, mid, step)
        assert np.array_equal(sdf_a, sdf_b) and np.array_equal(todo_a, todo_b)


def test_callback_octree_has_no_cpu_path():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    coords, _ = mesh_util.create_grid(8, 8, 8)
    with pytest.raises(Exception, match="CUDA|libpifu|no CPU"):
        mesh_util.eval_grid_octree(coords, lambda p: np.zeros(p.shape[1]), init_resolution=4)


def test_batch_eval_chunks_in_order():
    pts = np.arange(3 * 25, dtype=np.float64).reshape(3, 25)
    calls = []
    out = mesh_util.batch_eval(pts, lambda p: (calls.append(p.shape[1]), p[0] * 2)[1], num_samples=10)
    assert calls == [10, 10, 5] and np.array_equal(out, pts[0] * 2) and out.dtype == np.float64
    assert mesh_util.eval_grid(pts.reshape(3, 5, 5, 1), lambda p: p[1], 7).shape == (5, 5, 1)


def test_save_obj_format(tmp_path):
    v = np.array([[0.123456, -1.5, 2.0], [1, 2, 3]], dtype=np.float64)
    c = np.array([[0.5, 0.25, 1.0], [0, 0, 0]])
    f = np.array([[0, 1, 1]], dtype=np.int32)
    p = str(tmp_path / "m.obj")
    mesh_util.save_obj_mesh_with_color(p, v, f, c)
    lines = open(p).read().splitlines()
    assert lines[0] == "v 0.1235 -1.5000 2.0000 0.5000 0.2500 1.0000"
    assert lines[2] == "f 1 2 2"


def test_save_obj_bytes_identical_to_reference_loop(tmp_path):
    """The native writer reproduces printf('%.4f') exactly: random values, exact decimal ties
    (k/32 * 1e-3 style dyadic values), values just below / above a tie, signed zeros, huge values."""
    rng = np.random.default_rng(5)
    n = 20000
    v = rng.uniform(-3, 3, (n, 3))
    ties = np.array([0.03125, 0.09375, 0.00005, 0.00015, 0.12345, 2.5e-5, 1.00005, 0.99995, 9.99995, 8191.99995,
                     65536.00005, 123456.78905, 1e-300, -1e-9, 0.0, -0.0, 1e14, 3e15, 1e22, -7.5e15])
    near = np.concatenate([np.nextafter(ties, np.inf), np.nextafter(ties, -np.inf)])
    sp = np.concatenate([ties, near, -ties])
    v[:len(sp), 0] = sp
    v[:len(sp), 2] = sp[::-1]
    c = rng.uniform(0, 1, (n, 3))
    c[:len(sp), 1] = np.abs(sp) % 1.0
    f = rng.integers(0, n, (3 * n, 3)).astype(np.int32)
    f[0] = [2 ** 31 - 2, 0, 5]
    edges = [10 ** k + d for k in range(1, 10) for d in (-2, -1, 0)]   # every digit count of the 1-based index, both sides
    f[1:1 + len(edges) // 3] = np.array(edges, dtype=np.int64).reshape(-1, 3).astype(np.int32)
    p = str(tmp_path / "native.obj")
    mesh_util.save_obj_mesh_with_color(p, v, f, c)
    ref = io.StringIO()
    for idx, vv in enumerate(v):                       # the reference's loop, `mesh_util.py:192-197`
        cc = c[idx]
        ref.write('v %.4f %.4f %.4f %.4f %.4f %.4f\n' % (vv[0], vv[1], vv[2], cc[0], cc[1], cc[2]))
    for ff in f.astype(np.int64):
        ref.write('f %d %d %d\n' % (ff[0] + 1, ff[2] + 1, ff[1] + 1))
    assert open(p).read() == ref.getvalue()
    # empty mesh -> empty file; fewer colours than vertices is an error like the reference's IndexError
    mesh_util.save_obj_mesh_with_color(p, np.zeros((0, 3)), np.zeros((0, 3), np.int32), np.zeros((0, 3)))
    assert os.path.getsize(p) == 0
    with pytest.raises(IndexError):
        mesh_util.save_obj_mesh_with_color(p, v[:10], f[:1], c[:5])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver times beside ours): one JSON line with the contract's keys, the
    requested steps / warm-up, and the `config` object the GPU arm prints."""
    import json
    import subprocess
    import sys
    from helpers import ROOT
    import bench
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    res, what = bench.workload(1)
    assert d["config"] == bench.config_dict(1, res, what)


def test_train_mode_batchnorm_batches_are_refused():
    import torch
    from pifu_b200 import BasePIFuNet

    class M(torch.nn.Module):
        norm = "batch"
    m = M()
    BasePIFuNet.check_batch_statistics(m, 1)
    with pytest.raises(NotImplementedError):
        BasePIFuNet.check_batch_statistics(m, 2)
    m.eval()
    BasePIFuNet.check_batch_statistics(m, 2)


def test_save_obj_fixed4_property(tmp_path):
    """Any double - subnormals, huge magnitudes, both zeros, values a few ulps around every kind of decimal tie - prints as
    printf('%.4f') does (hypothesis draws the values; the library formats them as vertex coordinates)."""
    from hypothesis import given, settings, strategies as st
    path = str(tmp_path / "h.obj")
    anyf = st.floats(allow_nan=False, allow_infinity=False, width=64)
    near_tie = st.builds(lambda k, u, s: float(np.nextafter((2 * k + 1) / 20000.0 * s, np.inf if u > 0 else -np.inf)) if u else (2 * k + 1) / 20000.0 * s,
                         st.integers(0, 10 ** 9), st.integers(-1, 1), st.sampled_from([1.0, -1.0, 1e-3, 16.0, 4096.0]))

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.one_of(anyf, near_tie), min_size=6, max_size=600))
    def run(vals):
        n = len(vals) // 6
        a = np.array(vals[:6 * n], dtype=np.float64).reshape(n, 6)
        mesh_util.save_obj_mesh_with_color(path, a[:, :3].copy(), np.zeros((0, 3), np.int32), a[:, 3:].copy())
        want = "".join('v %.4f %.4f %.4f %.4f %.4f %.4f\n' % tuple(r) for r in a)
        assert open(path).read() == want
    run()


def _python_obj_loop(path):
    """The per-line loop the native reader replaced (the checker here): 'v ' lines -> floats, 'f ' lines -> 0-based
    (a, c, b) of the first three references."""
    vs, fs = [], []
    with open(path, newline="") as fh:
        for line in fh:
            if line.startswith("v "):
                vs.append([float(x) for x in line.split()[1:]])
            elif line.startswith("f "):
                a, b, c = (int(x.split("/")[0]) - 1 for x in line.split()[1:4])
                fs.append((a, c, b))
    return vs, fs


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_load_obj_native_reader_matches_line_loop(tmp_path, monkeypatch, threads):
    """pifu_obj_counts + pifu_read_obj against the Python loop, bit for bit: a written mesh (several segments per thread
    count), then a hand-made file with comments, vn / vt lines, a/b/c references, CRLF, exponents, signed zeros, long
    mantissas, no trailing newline."""
    monkeypatch.setenv("PIFU_OBJ_THREADS", threads)
    rng = np.random.default_rng(11)
    n = 30000
    v = rng.uniform(-300, 300, (n, 3)) * rng.choice([1.0, 1e-3, 1e3], (n, 1))
    c = rng.uniform(0, 1, (n, 3))
    f = rng.integers(0, n, (2 * n + 7, 3)).astype(np.int32)
    p = str(tmp_path / "w.obj")
    mesh_util.save_obj_mesh_with_color(p, v, f, c)
    V, F, C = mesh_util.load_obj_mesh_with_color(p)
    vs, fs = _python_obj_loop(p)
    assert np.array_equal(np.hstack([V, C]), np.array(vs)) and np.array_equal(F, np.array(fs, dtype=np.int32))
    assert np.array_equal(F, f)                                          # the order the writer was given
    q = str(tmp_path / "hand.obj")
    text = ("# a comment\r\nmtllib x.mtl\nv 1 2 3\nvn 0 0 1\nvt 0.5 0.5\n"
            "v -0.0 1e-3 2.5E+2\r\nv 0.1234567890123456789 -123456789012345678901234.5 .5\n"
            "v 7. +8 -9.000\nv inf -inf 1e400\n"
            "f 1/1/1 2/2/2 3/3/3\nf 4//1 5//2 1//3\r\nf 1 2 3 4\ng grp\nf 5 4 3")          # no trailing newline
    open(q, "w", newline="").write(text)
    V, F, C = mesh_util.load_obj_mesh_with_color(q)
    vs, fs = _python_obj_loop(q)
    assert C is None and V.shape == (5, 3) and F.shape == (4, 3)
    ref = np.array(vs)
    assert np.array_equal(V, ref) and np.array_equal(np.signbit(V), np.signbit(ref))
    assert np.array_equal(F, np.array(fs, dtype=np.int32))


def test_load_obj_native_reader_edges(tmp_path):
    from pifu_b200._lib import PifuError
    e = str(tmp_path / "empty.obj")
    open(e, "w").close()
    V, F, C = mesh_util.load_obj_mesh_with_color(e)
    assert V.shape == (0, 3) and F.shape == (0, 3) and C is None
    with pytest.raises(PifuError):
        mesh_util.load_obj_mesh_with_color(str(tmp_path / "missing.obj"))
    b = str(tmp_path / "bad.obj")
    open(b, "w").write("v 1 2 3\nv 1 x 3\nf 1 2 2\n")
    with pytest.raises(PifuError):
        mesh_util.load_obj_mesh_with_color(b)
    b2 = str(tmp_path / "bad2.obj")
    open(b2, "w").write("v 1 2 3 0 0 0\nv 1 2 3\nf 1 2\n")
    with pytest.raises(PifuError):
        mesh_util.load_obj_mesh_with_color(b2)
