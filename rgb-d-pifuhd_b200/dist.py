"""Multi-GPU sharding of the lattice: one process per GPU, every rank holds
the broadcast feature maps and weights, the lattice is split with no data-path collective, and
NCCL is used only to gather occupancy values.

* dense grid: contiguous id ranges = slabs along volume axis 0 (the lattice id is
  (i*R1 + j)*R2 + k), gathered to one rank;
* octree: every rank runs the (cheap, deterministic) frontier compaction / skip / fill
  bookkeeping on the full field, the level's frontier is cut into equal shares for the MLP
  evaluation - better balanced than slabs because the frontier hugs the surface - and the
  shares are all-gathered so all replicas stay bit-identical.

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
    return sdf32 if r == dst else None
