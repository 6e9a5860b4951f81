"""Slab form of the device octree (multi-GPU path) on ONE GPU: W "virtual ranks" - W library contexts on the
same device - walk the levels in lock step exactly as `dist.sharded_octree_slab` drives real ranks (coarse levels on the
local planes without exchange, fine levels through the concatenated frontier), with an analytic field as eval_func.
Every rank's planes [own - 1, own + 2) must equal the single-volume device octree bit for bit, and the slab meshes
must assemble into the whole-volume mesh."""
import pytest
import torch

from test_octree_mc_gpu import _analytic_on_device

pytestmark = pytest.mark.gpu


def _whole(eng, res, init, name):
    eng.octree_begin(res, init, 0.05)
    while True:
        step, ids = eng.octree_frontier()
        if step == 0:
            break
        eng.octree_commit(_analytic_on_device(name, ids, res))
    return eng.octree_export(want64=False, want32=True)[1].clone()


@pytest.mark.parametrize("name", ["ellipsoid", "ripple"])
@pytest.mark.parametrize("res,init,W", [(128, 16, 4), (128, 32, 3), (96, 12, 2), (256, 32, 8)])
def test_virtual_ranks_match_whole_volume(name, res, init, W):
    from pifu_b200 import dist as pdist, get_engine
    from pifu_b200.engine import Engine
    ref = _whole(get_engine("cuda"), res, init, name)
    engs = [Engine(0) for _ in range(W)]
    plan = [pdist.octree_slab_planes(res, init, W, r) for r in range(W)]          # (pb, pe, lb, le)
    live = [r for r in range(W) if plan[r][1] > plan[r][0]]
    for r in live:
        pb, pe, lb, le = plan[r]
        engs[r].octree_begin_slab(res, init, 0.05, lb, le, pb, pe)
    step = res // init
    plane = res * res
    evaluated = 0
    while step > 0:
        if W > 1 and step >= pdist.LOCAL_LEVEL_MIN_STEP:
            for r in live:
                pb, pe, lb, le = plan[r]
                engs[r].octree_set_frontier_planes(lb, le)
                ids = engs[r].octree_frontier()[1]
                evaluated += int(((ids >= pb * plane) & (ids < pe * plane)).sum())
                engs[r].octree_commit(_analytic_on_device(name, ids, res))
        else:
            parts = []
            for r in live:
                engs[r].octree_set_frontier_planes(plan[r][0], plan[r][1])
                parts.append(engs[r].octree_frontier()[1])
            flat = torch.cat(parts)
            assert bool((flat[1:] > flat[:-1]).all())                   # concatenation in rank order = C order of the volume
            evaluated += flat.numel()
            vals = _analytic_on_device(name, flat, res)
            for r in live:
                engs[r].octree_commit_pairs(flat, vals)
        step //= 2
    for r in live:
        pb, pe, lb, le = plan[r]
        field, first = engs[r].octree_field32()
        assert first == lb and field.shape[0] == le - lb
        lo, hi = max(pb - 1, 0), min(pe + 2, res)
        assert torch.equal(field[lo - lb:hi - lb], ref[lo:hi]), "rank %d planes [%d, %d)" % (r, lo, hi)
    # the evaluated set is the single-volume one (every point evaluated once, by its owner)
    eng0 = get_engine("cuda")
    eng0.octree_begin(res, init, 0.05)
    n_ref = 0
    while True:
        step, ids = eng0.octree_frontier()
        if step == 0:
            break
        n_ref += ids.numel()
        eng0.octree_commit(_analytic_on_device(name, ids, res))
    assert evaluated == n_ref
