"""CPU, gloo, world_size 2: the host-side sharding logic of the N>1 path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import orc

from pifu_b200 import dist as pdist


def test_shard_bounds_cover_and_align():
    for n, W, align in [(512 ** 3, 8, 512 * 512), (1000, 3, 1), (7, 8, 1), (64 ** 3, 2, 64 * 64), (0, 2, 1)]:
        prev = 0
        for r in range(W):
            b, e = pdist.shard_bounds(n, W, r, align)
            assert b == prev and e >= b
            assert b % align == 0 or b == n
            prev = e
        assert prev == n


class FakeEngine:
    """Host stand-in for the device octree state: analytic field, same stepwise protocol."""

    def __init__(self, res, fn):
        self.res, self.fn, self.device = res, fn, torch.device("cpu")

    def eval_grid(self, levels, res, calib, id_begin=0, id_end=None):
        ids = torch.arange(id_begin, id_end)
        return self.fn(ids)

    def octree_begin(self, res, init_resolution, threshold):
        from oracle import pifu_oracle
        self.sdf = np.zeros((res,) * 3)
        self.todo = np.zeros((res,) * 3, bool)
        self.todo[:-1, :-1, :-1] = True
        self.lat = np.zeros((res,) * 3, bool)
        self.step, self.thr, self.mu = res // init_resolution, threshold, pifu_oracle

    def octree_frontier(self):
        if self.step <= 0:
            return 0, torch.empty(0, dtype=torch.int64)
        self.lat[::self.step, ::self.step, ::self.step] = True
        self.test = self.lat & self.todo
        return self.step, torch.from_numpy(np.flatnonzero(self.test))

    def octree_commit(self, vals):
        self.sdf[self.test] = vals.numpy().astype(np.float64)
        self.todo[self.test] = False
        self._skip_fill()

    def _skip_fill(self):
        s = self.step
        if s > 1:
            v = self.sdf[::s, ::s, ::s]
            n = [k - 1 for k in v.shape]
            cs = [v[a:a + n[0], b:b + n[1], c:c + n[2]] for a in (0, 1) for b in (0, 1) for c in (0, 1)]
            lo, hi = np.minimum.reduce(cs), np.maximum.reduce(cs)
            centre = self.todo[s // 2::s, s // 2::s, s // 2::s][:n[0], :n[1], :n[2]]
            self.mu.octree_fill_gather(self.sdf, self.todo, ((hi - lo) < self.thr) & centre, 0.5 * (lo + hi), s)
        self.step = 0 if s <= 1 else s // 2

    def octree_export(self, want64=False, want32=True):
        return None, torch.from_numpy(self.sdf.astype(np.float32))


def _field(res):
    def fn(ids):
        k = ids % res
        j = (ids // res) % res
        i = ids // (res * res)
        x, y, z = (i.double() / res * 2 - 1), (j.double() / res * 2 - 1), (k.double() / res * 2 - 1)
        r = torch.sqrt((x / 0.35) ** 2 + (y / 0.8) ** 2 + (z / 0.3) ** 2)
        return torch.clamp(0.5 + 2.0 * (1.0 - r), 0, 1).float()
    return fn


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = 32
    eng = FakeEngine(res, _field(res))
    dense = pdist.sharded_eval_grid(eng, 2, res, None)
    stats = []
    octo = pdist.sharded_eval_grid_octree(eng, 2, res, None, init_resolution=8, threshold=0.05, stats=stats,
                                          evaluate=lambda ids: eng.fn(ids))
    if rank == 0:
        torch.save({"dense": dense, "octree": octo, "stats": stats}, out)
    else:
        assert dense is None and octo is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29571, out), nprocs=2, join=True)
    got = torch.load(out)
    res = 32
    fn = _field(res)
    ref_dense = fn(torch.arange(res ** 3)).view(res, res, res)
    assert torch.equal(got["dense"], ref_dense)
    coords = np.indices((res,) * 3).astype(np.float64)

    def eval_func(p):
        ids = (p[0] * res + p[1]) * res + p[2]
        return fn(torch.from_numpy(ids.astype(np.int64))).numpy()
    calls = []
    ref_oct = orc.eval_grid_octree(coords, lambda p: (calls.append(p.shape[1]), eval_func(p))[1],
                                   init_resolution=8, num_samples=10 ** 9)
    assert np.array_equal(got["octree"].numpy(), ref_oct.astype(np.float32))
    assert got["stats"] == calls


# ----------------------------------------------------------------------------- sharded mesh
def _oracle_slab(field, level, i0, g0, layers, ghost):
    """Stand-in for Engine.marching_cubes_slab: the CPU oracle's slab form, torch in / torch out."""
    from oracle import mc_oracle
    v, f, n, val, ng = mc_oracle.marching_cubes_slab(field.numpy(), level, i0, g0, layers, ghost)
    return torch.from_numpy(v), torch.from_numpy(f), torch.from_numpy(n), torch.from_numpy(val), ng


class FakeMeshEngine(FakeEngine):
    """+ the slab form of the device octree (octree.cu), restated in numpy on the local planes [lb, le): the global
    volume's last plane is never processed, the frontier covers the own planes only and carries global lattice ids,
    and a commit stores whichever (id, value) pairs fall inside the local planes."""
    marching_cubes_slab = staticmethod(_oracle_slab)

    def marching_cubes_slab_async(self, field, level, i0, g0, layers, ghost, cap=None):
        """The device engine's asynchronous form: outputs at a capacity, counts as a tensor; the first call is given a
        capacity that is too small, so the retry path of `sharded_marching_cubes` runs as well."""
        v, f, n, val, ng = _oracle_slab(field, level, i0, g0, layers, ghost)
        counts = torch.tensor([v.shape[0], f.shape[0], ng], dtype=torch.int64)
        cap = cap if cap is not None else (max(v.shape[0] - 3, 0), f.shape[0] + 5)
        if v.shape[0] > cap[0] or f.shape[0] > cap[1]:
            return (torch.zeros((cap[0], 3), dtype=torch.float64), torch.zeros((cap[1], 3), dtype=torch.int32),
                    torch.zeros((cap[0], 3)), torch.zeros(cap[0]), counts)
        return v, f, n, val, counts

    def eval_lattice_ids(self, levels, res, ids, calib):
        return self.fn(ids)

    def octree_begin_slab(self, res, init_resolution, threshold, lb, le, pb, pe):
        from oracle import pifu_oracle
        self.lb, self.own = lb, (pb - lb, pe - lb)
        self.sdf = np.zeros((le - lb, res, res))
        self.todo = np.zeros((le - lb, res, res), bool)
        self.todo[:, :-1, :-1] = True
        if le == res:
            self.todo[-1] = False
        self.lat = np.zeros((le - lb, res, res), bool)
        self.step, self.thr, self.mu = res // init_resolution, threshold, pifu_oracle
        self.slab = True

    def octree_frontier(self):
        if not getattr(self, "slab", False):
            return super().octree_frontier()
        if self.step <= 0:
            return 0, torch.empty(0, dtype=torch.int64)
        self.lat[::self.step, ::self.step, ::self.step] = True
        test = self.lat & self.todo
        test[:self.own[0]] = False
        test[self.own[1]:] = False
        self._last_ids = torch.from_numpy(np.flatnonzero(test) + self.lb * self.res * self.res)
        return self.step, self._last_ids

    def octree_set_frontier_planes(self, plane_begin, plane_end):
        self.own = (plane_begin - self.lb, plane_end - self.lb)

    def octree_commit(self, vals):
        if not getattr(self, "slab", False):
            return super().octree_commit(vals)
        self.octree_commit_pairs(self._last_ids, vals)

    def octree_commit_pairs(self, ids, vals):
        v = ids.numpy() - self.lb * self.res * self.res
        keep = (v >= 0) & (v < self.sdf.size)
        self.sdf.reshape(-1)[v[keep]] = vals.numpy()[keep].astype(np.float64)
        self.todo.reshape(-1)[v[keep]] = False
        self._skip_fill()

    def octree_field32(self):
        return torch.from_numpy(self.sdf.astype(np.float32)), self.lb


def _ripple(res):
    """A field whose skip decisions change from cell to cell (thin sheets, saddles): the slab margins are exercised."""
    def fn(ids):
        k = ids % res
        j = (ids // res) % res
        i = ids // (res * res)
        x, y, z = (i.double() / res * 2 - 1), (j.double() / res * 2 - 1), (k.double() / res * 2 - 1)
        v = 0.5 + 0.35 * torch.sin(5.1 * x + 1.3 * y * z) * torch.cos(3.7 * y - 2.2 * z) + 0.2 * (x * y - z * z)
        return torch.clamp(v, 0, 1).float()
    return fn


def _mesh_worker(rank, world, port, out, res, flat):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fn = (lambda ids: torch.full((ids.numel(),), 0.25)) if flat is True else (_ripple(res) if flat == "ripple" else _field(res))
    eng = FakeMeshEngine(res, fn)
    got = {}
    for mode in ("dense", "octree"):
        try:
            m = pdist.sharded_mesh(eng, 2, res, None, mode == "octree", level=0.5, init_resolution=8)
            got[mode] = m
            assert (m is None) == (rank != 0)
        except ValueError:
            got[mode] = "no surface"
    if rank == 0:
        torch.save(got, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,res,kind", [(2, 32, False), (3, 24, False), (4, 64, False), (4, 64, "ripple"), (3, 48, "ripple")])
def test_sharded_mesh_equals_whole_volume(tmp_path, world, res, kind):
    """Slab-by-slab extraction + fragment gather (ghost layer, halo planes, vertex renumbering)
    gives bit for bit the mesh of a sequential traversal of the whole volume - for the dense field and for the
    slab-sharded octree (own planes + margin, no boundary exchange: `dist.sharded_octree_slab`)."""
    from oracle import mc_oracle
    out = str(tmp_path / "mesh.pt")
    mp.spawn(_mesh_worker, args=(world, 29573 + world + (7 if kind else 0), out, res, kind), nprocs=world, join=True)
    got = torch.load(out)
    fn = _ripple(res) if kind == "ripple" else _field(res)
    dense = fn(torch.arange(res ** 3)).view(res, res, res).numpy()
    coords = np.indices((res,) * 3).astype(np.float64)
    octo = orc.eval_grid_octree(
        coords, lambda p: fn(torch.from_numpy(((p[0] * res + p[1]) * res + p[2]).astype(np.int64))).numpy(),
        init_resolution=8, num_samples=10 ** 9).astype(np.float32)
    for mode, vol in (("dense", dense), ("octree", octo)):
        rv, rf, rn, rval, _ = mc_oracle.marching_cubes(vol, 0.5)
        v, f, n, val = got[mode]
        assert np.array_equal(f.numpy(), rf) and np.array_equal(v.numpy(), rv)
        assert np.array_equal(n.numpy(), rn) and np.array_equal(val.numpy(), rval)


def test_sharded_mesh_without_surface_raises_everywhere(tmp_path):
    out = str(tmp_path / "mesh.pt")
    mp.spawn(_mesh_worker, args=(2, 29579, out, 16, True), nprocs=2, join=True)
    got = torch.load(out)
    assert got == {"dense": "no surface", "octree": "no surface"}
