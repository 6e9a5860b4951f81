"""CPU ORACLE (test infrastructure - NOT a product path).

Restatement, in functional torch-CPU fp32, of the reference's occupancy query path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  It is pinned against the reference
itself: ``oracle/make_golden.py`` imports the reference's own ``PIFuNetwNML`` /
``PIFuMRNet`` / ``mesh_util`` from ``/root/reference`` (authoring container only) and
stores their outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
file against those fixtures.

Each function cites the reference lines it follows.  The same library kernels the
reference runs on CPU (``baddbmm``, ``grid_sample``, ``conv1d``, ``group_norm``) are used
so that timing this port is a faithful stand-in for timing the reference on host cores.
"""
import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- projection
def orthogonal(points, calib):
    """`BasePIFuNet.py:25-43` (transform is None at every call site)."""
    rot = calib[:, :3, :3]
    trans = calib[:, :3, 3:4]
    return torch.baddbmm(trans, rot, points)


def perspective(points, calib):
    """`BasePIFuNet.py:45-65`."""
    rot = calib[:, :3, :3]
    trans = calib[:, :3, 3:4]
    homo = torch.baddbmm(trans, rot, points)
    xy = homo[:, :2, :] / homo[:, 2:3, :]
    return torch.cat([xy, homo[:, 2:3, :]], 1)


def project(points, calib, mode="orthogonal"):
    """`BasePIFuNet.py:79`: anything but the exact string 'orthogonal' selects perspective."""
    return orthogonal(points, calib) if mode == "orthogonal" else perspective(points, calib)


# --------------------------------------------------------------------------- sampling
def index(feat, uv):
    """`BasePIFuNet.py:11-23`: bilinear grid_sample, zeros padding, align_corners=True."""
    grid = uv.transpose(1, 2).unsqueeze(2)
    return F.grid_sample(feat, grid, mode="bilinear", padding_mode="zeros",
                         align_corners=True)[:, :, :, 0]


def index_closed_form(feat, uv):
    """Same op written out tap by tap (numpy fp32) - the formula the CUDA kernel implements.
    ix = (u+1)/2*(W-1), iy = (v+1)/2*(H-1); out-of-range taps contribute 0."""
    f = feat[0].numpy()
    C, H, W = f.shape
    u = uv[0, 0].numpy().astype(np.float32)
    v = uv[0, 1].numpy().astype(np.float32)
    ix = ((u + np.float32(1)) / np.float32(2)) * np.float32(W - 1)
    iy = ((v + np.float32(1)) / np.float32(2)) * np.float32(H - 1)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    fx = (ix - x0).astype(np.float32)
    fy = (iy - y0).astype(np.float32)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    out = np.zeros((C, u.shape[0]), np.float32)
    for dy, wy in ((0, np.float32(1) - fy), (1, fy)):
        for dx, wx in ((0, np.float32(1) - fx), (1, fx)):
            xx = x0 + dx
            yy = y0 + dy
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            w = np.where(ok, wx * wy, np.float32(0)).astype(np.float32)
            out += f[:, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * w[None]
    return torch.from_numpy(out)[None]


def depth_normalize(xyz, load_size, z_size):
    """`DepthNormalizer.py:17-25`: z * (loadSize // 2) / z_size, two fp32 roundings."""
    return xyz[:, 2:3, :] * (load_size // 2) / z_size


# --------------------------------------------------------------------------- MLP
def mlp_forward(feature, sd, n_layers, res_layers, merge_layer, norm="none", prefix=""):
    """`MLP.py:42-75`.  ``sd`` maps ``filters.{i}.weight|bias`` (+ ``norms.{i}.*``)."""
    y = feature
    tmpy = feature
    phi = None
    for i in range(n_layers):
        w = sd[prefix + "filters.%d.weight" % i]
        b = sd[prefix + "filters.%d.bias" % i]
        y = F.conv1d(y if i not in res_layers else torch.cat([y, tmpy], 1), w, b)
        if i != n_layers - 1:
            if norm == "group":
                y = F.group_norm(y, 32, sd[prefix + "norms.%d.weight" % i],
                                 sd[prefix + "norms.%d.bias" % i], 1e-5)
            elif norm == "batch_train":        # BatchNorm1d left in train mode (`reconstruction.py:288-289`): batch statistics
                y = F.batch_norm(y, None, None, sd[prefix + "norms.%d.weight" % i], sd[prefix + "norms.%d.bias" % i],
                                 True, 0.1, 1e-5)
            elif norm == "batch_eval":         # running statistics
                y = F.batch_norm(y, sd[prefix + "norms.%d.running_mean" % i], sd[prefix + "norms.%d.running_var" % i],
                                 sd[prefix + "norms.%d.weight" % i], sd[prefix + "norms.%d.bias" % i], False, 0.1, 1e-5)
            elif norm == "batch":
                raise NotImplementedError("say batch_train or batch_eval: the statistics depend on the module's mode")
            y = F.leaky_relu(y)
        if i == merge_layer:
            phi = y.clone()
    return torch.sigmoid(y), phi


def effective_merge_layer(merge_layer, filter_channels):
    """`MLP.py:25`."""
    return merge_layer if merge_layer > 0 else len(filter_channels) // 2


# --------------------------------------------------------------------------- nets
class CoarseState:
    """What `PIFuNetwNML.query` reads: MLP weights, the last feature map, option fields."""

    def __init__(self, sd, feat, opt, mode="orthogonal"):
        self.sd, self.feat, self.opt, self.mode = sd, feat, opt, mode
        self.n_layers = len(opt.mlp_dim) - 1
        self.merge = effective_merge_layer(opt.merge_layer, opt.mlp_dim)


class FineState:
    """What `PIFuMRNet.query` reads (fine MLP is built with merge_layer=-1, `PIFuMRNet.py:41-45`)."""

    def __init__(self, sd, feat, opt, coarse, mode="orthogonal"):
        self.sd, self.feat, self.opt, self.coarse, self.mode = sd, feat, opt, coarse, mode
        self.n_layers = len(opt.mlp_dim) - 1
        self.merge = effective_merge_layer(-1, opt.mlp_dim)


def query_coarse(st, points, calib):
    """`PIFuNetwNML.py:99-141` with one feature map (eval mode, `:96-97`).
    Returns (preds [B,1,N], phi [B,C,N])."""
    xyz = project(points, calib, st.mode)
    xy = xyz[:, :2, :]
    inb = (xyz >= -1) & (xyz <= 1)
    inb = (inb[:, 0, :] & inb[:, 1, :] & inb[:, 2, :])[:, None, :].float()
    sp = depth_normalize(xyz, st.opt.loadSize, st.opt.z_size)
    feat = torch.cat([index(st.feat, xy), sp], 1)
    pred, phi = mlp_forward(feat, st.sd, st.n_layers, st.opt.mlp_res_layers, st.merge,
                            st.opt.mlp_norm)
    return inb * pred, phi


def query_fine(st, points, calib_local, calib_global=None):
    """`PIFuMRNet.py:119-186`, single-level call form (`:131-137`) and B2 == 1.
    Returns (preds [B,1,N], preds_low [1,B,1,N], phi)."""
    if calib_global is None:
        calib_global = calib_local
    xyz = project(points, calib_local, st.mode)
    xy = xyz[:, :2, :]
    inb = (xyz >= -1) & (xyz <= 1)
    inb = (inb[:, 0, :] & inb[:, 1, :])[:, None, :].float()
    low, phi = query_coarse(st.coarse, points, calib_global)
    feat = torch.cat([index(st.feat, xy), phi], 1)
    pred = mlp_forward(feat, st.sd, st.n_layers, st.opt.mlp_res_layers, st.merge,
                       st.opt.mlp_norm)[0]
    return inb * pred, low[None], phi


def calc_normal_fine(st, points, calib_local, calib_global, delta=0.001, return_raw=False):
    """`PIFuMRNet.py:188-243`, B2 == 1, forward differences.  points [B,3,N] -> nml [B,3,N]."""
    pts = [points.clone() for _ in range(4)]
    for a in range(3):
        pts[a + 1][:, a, :] += delta
    pall = torch.stack(pts, 3).view(points.shape[0], 3, -1)
    xyz = project(pall, calib_local, st.mode)
    _, phi = query_coarse(st.coarse, pall, calib_global)
    feat = torch.cat([index(st.feat, xyz[:, :2, :]), phi], 1)
    pred = mlp_forward(feat, st.sd, st.n_layers, st.opt.mlp_res_layers, st.merge,
                       st.opt.mlp_norm)[0]
    pred = pred.view(pred.shape[0], pred.shape[1], -1, 4)
    d = [pred[:, :, :, a + 1] - pred[:, :, :, 0] for a in range(3)]
    raw = -torch.cat(d, 1)
    nml = F.normalize(raw, dim=1, eps=1e-8)
    return (nml, raw) if return_raw else nml


# --------------------------------------------------------------------------- lattice
def lattice_coords(res, calib):
    """`mesh_util.py:12-38` + `:59-65`: float64 lattice b_min + idx*(2/res), then
    [p,1] @ inv(calib)^T.  Returns (coords [3,R,R,R] f64, mat [4,4] f64, calib_inv)."""
    rx = ry = rz = res
    idx = np.mgrid[:rx, :ry, :rz].reshape(3, -1)
    mat = np.eye(4)
    mat[0, 0], mat[1, 1], mat[2, 2] = 2.0 / rx, 2.0 / ry, 2.0 / rz
    mat[0:3, 3] = -1.0
    coords = np.matmul(mat[:3, :3], idx) + mat[:3, 3:4]
    calib_inv = np.linalg.inv(calib[0].cpu().numpy())
    homog = np.concatenate([coords.T, np.ones((coords.shape[1], 1))], 1)
    coords = np.matmul(homog, calib_inv.T)[:, :3].T.reshape(3, rx, ry, rz)
    return coords, mat, calib_inv


def batch_eval(points, eval_func, num_samples):
    """`mesh_util.py:98-114`: in-order chunks into a float64 result."""
    n = points.shape[1]
    out = np.zeros(n)
    for s in range(0, n, num_samples):
        out[s:s + num_samples] = eval_func(points[:, s:s + num_samples])
    return out


def eval_grid(coords, eval_func, num_samples):
    """`mesh_util.py:116-120`."""
    shape = coords.shape[1:4]
    return batch_eval(coords.reshape(3, -1), eval_func, num_samples).reshape(shape)


def make_eval_func(query, calib):
    """`mesh_util.py:67-74`: float64 points -> float32 tensor -> query -> preds[0][0] numpy."""
    def eval_func(points):
        samples = torch.from_numpy(np.expand_dims(points, 0)).float()
        return query(samples, calib)[0][0].detach().numpy()
    return eval_func


# --------------------------------------------------------------------------- octree
def eval_grid_octree(coords, eval_func, init_resolution=64, threshold=0.05,
                     num_samples=512 * 512 * 512, stats=None):
    """`mesh_util.py:124-187`, sequential semantics: float64 field, last planes never
    evaluated (`:135`), inclusive fill range, later skip cells overwrite earlier ones."""
    res = coords.shape[1:4]
    sdf = np.zeros(res)
    todo = np.zeros(res, dtype=bool)
    todo[:-1, :-1, :-1] = True
    lattice = np.zeros(res, dtype=bool)
    step = res[0] // init_resolution
    while step > 0:
        lattice[0:res[0]:step, 0:res[1]:step, 0:res[2]:step] = True
        test = lattice & todo
        if stats is not None:
            stats.append((step, int(test.sum())))
        sdf[test] = batch_eval(coords[:, test], eval_func, num_samples)
        todo[test] = False
        if step <= 1:
            break
        gx, gy, gz = (np.arange(0, r, step) for r in res)
        v = sdf[np.ix_(gx, gy, gz)]
        corners = np.stack([v[a:v.shape[0] - 1 + a, b:v.shape[1] - 1 + b, c:v.shape[2] - 1 + c]
                            for a in (0, 1) for b in (0, 1) for c in (0, 1)], 0)
        lo = corners.min(0)
        hi = corners.max(0)
        mid = 0.5 * (lo + hi)
        centre = todo[np.ix_(gx[:-1] + step // 2, gy[:-1] + step // 2, gz[:-1] + step // 2)]
        skip = ((hi - lo) < threshold) & centre
        for cx, cy, cz in zip(*np.where(skip)):
            x, y, z = cx * step, cy * step, cz * step
            sdf[x:x + step + 1, y:y + step + 1, z:z + step + 1] = mid[cx, cy, cz]
            todo[x:x + step + 1, y:y + step + 1, z:z + step + 1] = False
        step //= 2
    return sdf


def octree_fill_gather(sdf, todo, skip, mid, step):
    """Data-parallel restatement of the fill loop (`mesh_util.py:181-184`) used by the CUDA
    kernel: voxel p is covered, per axis, by cell p//step (if it exists) and - only when
    p % step == 0 - by cell p//step - 1; among covering skip cells the lexicographically
    largest (x, y, z) wrote last.  Mutates and returns (sdf, todo)."""
    R = sdf.shape
    nc = skip.shape
    p = [np.arange(r) for r in R]
    best_val = np.zeros(R)
    found = np.zeros(R, dtype=bool)
    # visit candidates in descending lexicographic order: hi cell first on each axis
    for ox in (0, 1):
        for oy in (0, 1):
            for oz in (0, 1):
                cand = []
                ok_axes = []
                for ax, o in zip(range(3), (ox, oy, oz)):
                    c = p[ax] // step - o
                    ok = (c >= 0) & (c < nc[ax])
                    if o == 1:
                        ok &= (p[ax] % step == 0)
                    cand.append(np.clip(c, 0, max(nc[ax] - 1, 0)))
                    ok_axes.append(ok)
                ok = ok_axes[0][:, None, None] & ok_axes[1][None, :, None] & ok_axes[2][None, None, :]
                s = skip[np.ix_(*cand)] & ok & ~found
                best_val = np.where(s, mid[np.ix_(*cand)], best_val)
                found |= s
    sdf[found] = best_val[found]
    todo[found] = False
    return sdf, todo
