"""Multi-GPU sharding of the lattice: one process per GPU, every rank holds
the broadcast feature maps and weights, the lattice is split with no data-path collective, and
NCCL is used only to gather occupancy values.

* dense grid: contiguous id ranges = slabs along volume axis 0 (the lattice id is
  (i*R1 + j)*R2 + k), gathered to one rank;
* octree, mesh path (`sharded_mesh`): the bookkeeping is sharded by slab too.  A rank keeps its own planes plus a
  margin of twice the initial stride, compacts the frontier of its own planes, the ranks all-gather the frontier ids,
  evaluate equal shares of the concatenated list (better balanced than slabs - the frontier hugs the surface),
  all-gather the values, and every rank commits the pairs that fall inside its planes.  The margin makes any
  boundary-plane exchange unnecessary (octree.cu) and already holds the two halo planes slab marching cubes reads;
* octree, field path (`sharded_eval_grid_octree`, when a caller wants the whole field on one rank): every rank runs
  the bookkeeping on the full field and only the MLP evaluation is split.

The helpers are backend-agnostic (gloo on CPU in the tests, NCCL on the box)."""
import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def shard_bounds(n, world, r, align=1):
    """Contiguous share [begin, end) of `n` items for rank r; shares are multiples of `align`
    (a whole number of lattice planes for the dense grid) except possibly the last."""
    units = (n + align - 1) // align
    per = (units + world - 1) // world
    b = min(n, r * per * align)
    e = min(n, (r + 1) * per * align)
    return b, e


def gather_concat(local, n_total, per_rank, group=None, dst=None):
    """Concatenate the ranks' shares (each at most `per_rank` long, in rank order) into a
    [n_total] tensor.  dst=None: all ranks receive it (all_gather); otherwise only rank dst."""
    W = world_size(group)
    if W == 1:
        return local[:n_total]
    pad = torch.zeros(per_rank, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    if dst is None:
        buf = torch.empty(W * per_rank, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(buf, pad, group=group)
        return buf[:n_total]
    if rank(group) == dst:
        parts = [torch.empty_like(pad) for _ in range(W)]
        dist.gather(pad, parts, dst=dst, group=group)
        return torch.cat(parts)[:n_total]
    dist.gather(pad, None, dst=dst, group=group)
    return None


def gather_varlen(local, counts, group=None, dst=0):
    """Concatenate per-rank tensors whose leading sizes `counts` differ, on rank `dst` (None elsewhere)."""
    W = world_size(group)
    if W == 1:
        return local
    per = max(max(counts), 1)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    if rank(group) == dst:
        parts = [torch.empty_like(pad) for _ in range(W)]
        dist.gather(pad, parts, dst=dst, group=group)
        return torch.cat([p[:n] for p, n in zip(parts, counts)], 0)
    dist.gather(pad, None, dst=dst, group=group)
    return None


def exchange_halo(local, group=None):
    """local: this rank's planes [n, R1, R2] (n >= 2).  Returns (below, above): the previous rank's
    last plane [1, R1, R2] (None on rank 0) and the next rank's first two planes (None on the last
    rank) - what slab marching cubes needs beyond its own planes (~1 MB at 512^2)."""
    W, r = world_size(group), rank(group)
    if W == 1:
        return None, None
    if local.shape[0] < 2:        # rank-local: callers validate the cut on every rank before any collective (sharded_mesh)
        raise ValueError("slab marching cubes needs at least two planes per rank")
    edge = torch.stack([local[0], local[1], local[-1]], 0).contiguous()
    buf = torch.empty((W * 3,) + tuple(edge.shape[1:]), dtype=edge.dtype, device=edge.device)
    dist.all_gather_into_tensor(buf, edge, group=group)
    buf = buf.view((W,) + tuple(edge.shape))
    below = buf[r - 1, 2:3] if r > 0 else None
    above = buf[r + 1, 0:2] if r < W - 1 else None
    return below, above


def _pack_fragment(v, f, n, val):
    """One byte buffer per rank: [verts f64 | normals f32 | values f32 | faces i32] (one gather instead of four)."""
    def raw(t):
        t = t.contiguous().reshape(-1)
        return t.view(torch.uint8) if t.numel() else torch.empty(0, dtype=torch.uint8, device=t.device)
    return torch.cat([raw(v), raw(n), raw(val), raw(f)])


def _unpack_fragment(buf, nv, nf):
    def typed(o, count, dtype, shape):
        if count == 0:
            return torch.empty(shape, dtype=dtype, device=buf.device)
        return buf[o:o + count * dtype.itemsize].view(dtype).view(shape)
    v = typed(0, nv * 3, torch.float64, (nv, 3))
    n = typed(nv * 24, nv * 3, torch.float32, (nv, 3))
    val = typed(nv * 36, nv, torch.float32, (nv,))
    f = typed(nv * 40, nf * 3, torch.int32, (nf, 3))
    return v, f, n, val


def sharded_marching_cubes(mc_slab, planes, level, R0, pb, pe, group=None, dst=0, mc_async=None, note_counts=None):
    """Marching cubes of a volume sharded along axis 0; every rank extracts the cells of its own
    planes [pb, pe) and the fragments are gathered (vertex numbers offset by the slabs before it)
    into exactly the mesh a traversal of the whole volume produces.

    planes(lo, hi) -> float32 [hi - lo, R1, R2] device tensor of global planes [lo, hi);
    mc_slab = Engine.marching_cubes_slab (or a stand-in with the same signature).
    Returns (verts, faces, normals, values) tensors on `dst`, None elsewhere."""
    W = world_size(group)
    cells_end = min(pe, R0 - 1)
    if W > 1 and mc_async is not None:
        # device path: the extraction is queued without a host synchronisation; the counts of all ranks come back with ONE
        # read (all-gather of three numbers), and the fragments travel in one packed gather
        dev = planes(pb, pb).device
        have = cells_end > pb
        if have:
            lo, hi = max(pb - 1, 0), min(pe + 2, R0)
            sub = planes(lo, hi)
            v, f, n, val, cnt = mc_async(sub, level, lo, R0, cells_end - lo, pb > 0)
        else:
            cnt = torch.zeros(3, device=dev, dtype=torch.int64)
        allc = torch.empty(W * 3, device=dev, dtype=torch.int64)
        dist.all_gather_into_tensor(allc, cnt, group=group)
        allc = allc.view(W, 3).tolist()
        r = rank(group)
        tv, tf, ng = (int(x) for x in allc[r])
        if have and (tv > v.shape[0] or tf > f.shape[0]):          # capacity exceeded: nothing was written, extract again
            v, f, n, val, cnt = mc_async(sub, level, lo, R0, cells_end - lo, pb > 0, cap=(tv, tf))
        if have and note_counts is not None:
            note_counts(tv, tf)
        nvs = [int(c[0] - c[2]) for c in allc]
        nfs = [int(c[1]) for c in allc]
        if sum(nvs) == 0:
            raise ValueError("No surface found at the given iso value (or level outside the data range)")
        first = sum(nvs[:r])
        if have:
            frag = _pack_fragment(v[ng:tv], f[:tf] + (first - ng), n[ng:tv], val[ng:tv])
        else:
            frag = torch.empty(0, device=dev, dtype=torch.uint8)
        sizes = [a * 40 + b * 12 for a, b in zip(nvs, nfs)]
        per = max(max(sizes), 1)
        pad = torch.empty(per, device=dev, dtype=torch.uint8)
        pad[:frag.numel()] = frag
        if r == dst:
            parts = [torch.empty_like(pad) for _ in range(W)]
            dist.gather(pad, parts, dst=dst, group=group)
            pieces = [_unpack_fragment(p_, a, b) for p_, a, b in zip(parts, nvs, nfs)]
            return tuple(torch.cat([pc[i] for pc in pieces], 0) for i in range(4))
        dist.gather(pad, None, dst=dst, group=group)
        return None
    if cells_end > pb:
        lo, hi = max(pb - 1, 0), min(pe + 2, R0)
        ghost = pb > 0
        v, f, n, val, ng = mc_slab(planes(lo, hi), level, lo, R0, cells_end - lo, ghost)
        v, n, val = v[ng:], n[ng:], val[ng:]
    else:
        ref = planes(pb, pb)
        v = torch.empty((0, 3), dtype=torch.float64, device=ref.device)
        f = torch.empty((0, 3), dtype=torch.int32, device=ref.device)
        n = torch.empty((0, 3), dtype=torch.float32, device=ref.device)
        val = torch.empty((0,), dtype=torch.float32, device=ref.device)
        ng = 0
    if W == 1:
        if v.shape[0] == 0:
            raise ValueError("No surface found at the given iso value (or level outside the data range)")
        return v, f, n, val
    cnt = torch.tensor([v.shape[0], f.shape[0]], dtype=torch.int64, device=v.device)
    allc = torch.empty(W * 2, dtype=torch.int64, device=v.device)
    dist.all_gather_into_tensor(allc, cnt, group=group)
    allc = allc.view(W, 2).cpu()
    nvs, nfs = [int(x) for x in allc[:, 0]], [int(x) for x in allc[:, 1]]
    if sum(nvs) == 0:
        raise ValueError("No surface found at the given iso value (or level outside the data range)")
    first = sum(nvs[:rank(group)])                       # global number of this slab's first own vertex
    f = f + (first - ng)
    out = [gather_varlen(t, c, group=group, dst=dst) for t, c in ((v, nvs), (f, nfs), (n, nvs), (val, nvs))]
    return tuple(out) if rank(group) == dst else None


def sharded_eval_grid(eng, levels, res, calib, group=None, dst=0):
    """Dense lattice, slab-sharded along axis 0.  Returns the [R,R,R] float32 field on `dst`
    (None on the other ranks)."""
    W, r = world_size(group), rank(group)
    total = res * res * res
    plane = res * res
    b, e = shard_bounds(total, W, r, align=plane)
    per = shard_bounds(total, W, 0, align=plane)[1]
    local = eng.eval_grid(levels, res, calib, id_begin=b, id_end=e) if e > b else \
        torch.empty(0, device=eng.device, dtype=torch.float32)
    full = gather_concat(local, total, per, group=group, dst=dst)
    return None if full is None else full.view(res, res, res)


def sharded_eval_grid_octree(eng, levels, res, calib, init_resolution=64, threshold=0.05, group=None,
                             dst=0, stats=None, evaluate=None):
    """Device octree with each level's frontier split evenly across ranks.  `evaluate(ids)`
    defaults to the engine's lattice-id query.  Returns the float32 field on `dst`."""
    W, r = world_size(group), rank(group)
    if evaluate is None:
        def evaluate(ids):
            return eng.eval_lattice_ids(levels, res, ids, calib)
    eng.octree_begin(res, init_resolution, threshold)
    while True:
        step, ids = eng.octree_frontier()
        if step == 0:
            break
        n = ids.numel()
        if stats is not None:
            stats.append(n)
        b, e = shard_bounds(n, W, r)
        per = shard_bounds(n, W, 0)[1]
        mine = evaluate(ids[b:e]) if e > b else torch.empty(0, device=ids.device, dtype=torch.float32)
        eng.octree_commit(gather_concat(mine, n, per, group=group, dst=None) if n else mine)
    _, sdf32 = eng.octree_export(want64=False, want32=True)
    return sdf32 if (dst is None or r == dst) else None


def octree_slab_planes(res, init_resolution, W, r):
    """(own_begin, own_end, plane_begin, plane_end) of rank r: own planes are whole multiples of the initial stride, the
    bookkeeping margin is twice that stride on either side (octree.cu)."""
    s0 = max(res // init_resolution, 1)
    pb, pe = shard_bounds(res, W, r, align=s0)
    return pb, pe, max(0, pb - 2 * s0), min(res, pe + 2 * s0)


LOCAL_LEVEL_MIN_STEP = 4      # levels of stride >= this are evaluated by every rank on its own local planes, without communication


def sharded_octree_slab(eng, levels, res, calib, init_resolution=64, threshold=0.05, group=None, stats=None,
                        evaluate=None):
    """Slab-sharded device octree.  Returns (field, plane_begin, own_begin, own_end): the float32 field of this rank's
    planes [plane_begin, plane_begin + field.shape[0]) - valid on [own_begin - 1, own_end + 2) - or field None for a
    rank that owns no plane.

    Coarse levels (stride >= LOCAL_LEVEL_MIN_STEP: a few hundred thousand points at 512^3) run without any collective:
    every rank compacts and evaluates the frontier of ALL its local planes, margin included - 1.5 x its fair share of a
    small level, against three collectives and two host synchronisations saved.  The values are the ones the neighbours
    compute for the same points (a point's value does not depend on the call it is evaluated in), and inside the region
    that matters the frontier is the true one because the bookkeeping state is (octree.cu).  The fine levels, which
    hold 90 % of the points, are balanced: ids all-gathered, equal shares evaluated, values all-gathered."""
    W, r = world_size(group), rank(group)
    if evaluate is None:
        def evaluate(ids):
            return eng.eval_lattice_ids(levels, res, ids, calib)
    pb, pe, lb, le = octree_slab_planes(res, init_resolution, W, r)
    mine = pe > pb
    dev = eng.device
    if mine:
        eng.octree_begin_slab(res, init_resolution, threshold, lb, le, pb, pe)
    step = res // init_resolution
    plane = res * res
    while step > 0:                                   # every rank walks the same levels, in lock step
        if W > 1 and step >= LOCAL_LEVEL_MIN_STEP:
            if mine:
                eng.octree_set_frontier_planes(lb, le)
                ids = eng.octree_frontier()[1]
                eng.octree_commit(evaluate(ids) if ids.numel() else torch.empty(0, device=dev, dtype=torch.float32))
            if stats is not None:                     # points evaluated by their owners = the single-device count
                own = ((ids >= pb * plane) & (ids < pe * plane)).sum().reshape(1) if mine else torch.zeros(1, device=dev, dtype=torch.int64)
                dist.all_reduce(own, group=group)
                stats.append(int(own.item()))
            step //= 2
            continue
        if mine:
            eng.octree_set_frontier_planes(pb, pe)
        ids = eng.octree_frontier()[1] if mine else torch.empty(0, device=dev, dtype=torch.int64)
        n = int(ids.numel())
        cnt = torch.tensor([n], device=dev, dtype=torch.int64)
        allc = torch.empty(W, device=dev, dtype=torch.int64)
        if W > 1:
            dist.all_gather_into_tensor(allc, cnt, group=group)
        else:
            allc = cnt
        counts = [int(x) for x in allc.tolist()]
        total = sum(counts)
        if stats is not None:
            stats.append(total)
        if total:
            per = max(counts)
            if W > 1:
                pad = torch.full((per,), -1, device=dev, dtype=torch.int64)
                pad[:n] = ids
                buf = torch.empty(W * per, device=dev, dtype=torch.int64)
                dist.all_gather_into_tensor(buf, pad, group=group)
                flat = torch.cat([buf[q * per:q * per + counts[q]] for q in range(W)])
            else:
                flat = ids
            b, e = shard_bounds(total, W, r)
            share = evaluate(flat[b:e]) if e > b else torch.empty(0, device=dev, dtype=torch.float32)
            vals = gather_concat(share, total, shard_bounds(total, W, 0)[1], group=group, dst=None)
        else:
            flat = torch.empty(0, device=dev, dtype=torch.int64)
            vals = torch.empty(0, device=dev, dtype=torch.float32)
        if mine:
            eng.octree_commit_pairs(flat, vals)
        step //= 2
    if not mine:
        return None, lb, pb, pe
    field, first = eng.octree_field32()
    return field, first, pb, pe


def sharded_mesh(eng, levels, res, calib, use_octree, level=0.5, init_resolution=64, threshold=0.05,
                 group=None, dst=0, stats=None):
    """Field + iso-surface of a res^3 lattice over all ranks (north_star: z-slab sharding, NCCL only
    for the halo planes / frontier values and the mesh fragments).  Dense: every rank evaluates and
    extracts its own slab, one halo exchange of three planes.  Octree: the replicated bookkeeping
    leaves the whole field on every rank, which then extracts its own slab.  Returns the mesh
    tensors on `dst` (None elsewhere); raises ValueError on every rank when there is no surface.
    (Octree: slab-sharded bookkeeping, `sharded_octree_slab`.)"""
    W, r = world_size(group), rank(group)
    plane = res * res
    b, e = shard_bounds(res * plane, W, r, align=plane)
    pb, pe = b // plane, e // plane
    if W > 1 and not use_octree:
        # every rank checks every rank's share (shard_bounds is deterministic), so all of them raise together and
        # none is left waiting in the halo exchange
        for q in range(W):
            bq, eq = shard_bounds(res * plane, W, q, align=plane)
            if (eq - bq) // plane < 2:
                raise ValueError("a %d^3 lattice cut over %d ranks leaves rank %d fewer than the two planes slab marching "
                                 "cubes needs" % (res, W, q))
    if use_octree and res // init_resolution > 0:
        field, first, pb, pe = sharded_octree_slab(eng, levels, res, calib, init_resolution, threshold, group=group,
                                                   stats=stats)

        def planes(lo, hi):
            if field is None:
                return torch.empty((0, res, res), device=eng.device, dtype=torch.float32)
            return field[lo - first:hi - first]
    elif use_octree:
        # resolution < init_resolution: the reference's loop never runs, the field is all zero (`mesh_util.py:138`)
        field = sharded_eval_grid_octree(eng, levels, res, calib, init_resolution, threshold, group=group,
                                         dst=None, stats=stats)

        def planes(lo, hi):
            return field[lo:hi]
    else:
        local = eng.eval_grid(levels, res, calib, id_begin=b, id_end=e).view(pe - pb, res, res) if e > b else \
            torch.empty((0, res, res), device=eng.device, dtype=torch.float32)
        below, above = exchange_halo(local, group=group)

        def planes(lo, hi):
            parts = []
            if lo < pb:
                parts.append(below[below.shape[0] - (pb - lo):])
            parts.append(local[max(lo - pb, 0):max(min(hi, pe) - pb, 0)])
            if hi > pe:
                parts.append(above[:hi - pe])
            return torch.cat(parts, 0) if len(parts) > 1 else parts[0]
    return sharded_marching_cubes(eng.marching_cubes_slab, planes, level, res, pb, pe, group=group, dst=dst,
                                  mc_async=getattr(eng, "marching_cubes_slab_async", None),
                                  note_counts=getattr(eng, "marching_cubes_note_counts", None))
