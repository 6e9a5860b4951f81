"""Thin Python face of one `pifu_ctx` (one per CUDA device).

PyTorch is plumbing here: it owns device memory and the stream; every kernel is in
libpifu_b200.so.  The engine snapshots MLP weights and feature maps of the live
``nn.Module``s and re-snapshots when their tensors change (``load_state_dict``,
a new ``filter*`` call)."""
import ctypes

import numpy as np
import torch

from . import _lib

_engines = {}


def get_engine(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.PifuError("pifu_b200 runs on CUDA devices only (got %s); there is no CPU path" % device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(idx)
    return _engines[idx]


def _stream(device_index):
    return ctypes.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)


def _calib16(calib):
    """[4,4] (or [3,4]) tensor/array -> ctypes float[16] row-major."""
    c = torch.as_tensor(calib).detach().to("cpu", torch.float32).reshape(-1, 4)
    full = torch.eye(4, dtype=torch.float32)
    full[: c.shape[0]] = c
    return (ctypes.c_float * 16)(*full.reshape(-1).tolist()), full


class Engine:
    def __init__(self, device_index):
        self.lib = _lib.load()
        self.device_index = device_index
        self.device = torch.device("cuda", device_index)
        h = ctypes.c_void_p()
        _lib.check(self.lib.pifu_create(device_index, ctypes.byref(h)))
        self.h = h
        self._mlp_key = [None, None]
        self._feat_key = [None, None]
        self._opt_key = None
        self._normalised = [False, False]
        self._keep = {}
        self._hold = {}
        self._mc_hint = {}
        self._copy_stream = None
        self.precision = 0
        import os
        if os.environ.get("PIFU_PRECISION"):
            self.set_precision(os.environ["PIFU_PRECISION"])

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.pifu_destroy(self.h)
                self.h = None
        except Exception:  # noqa: BLE001
            pass

    # ------------------------------------------------------------------ snapshots
    @staticmethod
    def _tensor_key(ts):
        """Identity of live tensors.  Only meaningful while the caller keeps `ts` referenced (see `_hold`): a freed
        tensor's id(), address and version count can all be re-issued to a different tensor."""
        return tuple((id(t), t.data_ptr(), t._version, tuple(t.shape)) for t in ts)

    def sync_mlp(self, level, mlp, owner_id=None):
        """mlp: object with .filters (Conv1d list), .norms, .res_layers, .merge_layer, .norm, .filter_channels.

        mlp_norm (`MLP.py:36-41`): 'group' -> GroupNorm(32) statistics over the points of each call;
        'batch' in train mode -> batch statistics per channel; 'batch' in eval mode is a per-channel
        affine map of the running statistics and is folded into the Conv1d weights here."""
        norm = mlp.norm if mlp.norm in ("batch", "group") else "none"
        norms = list(mlp.norms) if norm != "none" else []
        fold = norm == "batch" and not mlp.training
        params = []
        for f in mlp.filters:
            params += [f.weight, f.bias]
        for nm in norms:
            params += [nm.weight, nm.bias]
            if fold:
                params += [nm.running_mean, nm.running_var]
        key = (norm, fold, self._tensor_key(params))
        if self._mlp_key[level] == key:
            return
        # the keyed tensors stay referenced for as long as their key is cached, so no later net can be handed
        # the same (id, address, version) triple
        self._hold[("mlp", level)] = params
        if level == 0:
            self._mlp_key[1] = None
        ws, bs = [], []
        for i, f in enumerate(mlp.filters):
            w = f.weight.detach().to(self.device, torch.float32).reshape(f.weight.shape[0], -1)
            b = f.bias.detach().to(self.device, torch.float32)
            if fold and i < len(norms):
                nm = norms[i]
                scale = (nm.weight.detach().double() / torch.sqrt(nm.running_var.detach().double() + nm.eps)).to(self.device)
                shift = nm.bias.detach().double().to(self.device) - nm.running_mean.detach().double().to(self.device) * scale
                w = (w.double() * scale[:, None]).float()
                b = (b.double() * scale + shift).float()
            ws.append(w.contiguous())
            bs.append(b.contiguous())
        ch = list(mlp.filter_channels)
        res = list(mlp.res_layers)
        n = len(ws)
        wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in ws])
        bp = (ctypes.c_void_p * n)(*[b.data_ptr() for b in bs])
        _lib.check(self.lib.pifu_set_mlp(self.h, level, len(ch), (ctypes.c_int * len(ch))(*ch), len(res),
                                         (ctypes.c_int * max(len(res), 1))(*(res or [0])),
                                         int(mlp.merge_layer), wp, bp, _stream(self.device_index)))
        if norm != "none" and not fold:
            gs = [nm.weight.detach().to(self.device, torch.float32).contiguous() for nm in norms]
            be = [nm.bias.detach().to(self.device, torch.float32).contiguous() for nm in norms]
            gp = (ctypes.c_void_p * len(gs))(*[t.data_ptr() for t in gs])
            bp2 = (ctypes.c_void_p * len(be))(*[t.data_ptr() for t in be])
            groups = norms[0].num_groups if norm == "group" else 0
            _lib.check(self.lib.pifu_set_mlp_norm(self.h, level, int(groups), float(norms[0].eps), gp, bp2,
                                                  _stream(self.device_index)))
        torch.cuda.current_stream(self.device_index).synchronize()
        self._mlp_key[level] = key
        self._normalised[level] = norm != "none" and not fold

    def normalised(self, levels):
        """True when a level's MLP carries call-wide statistics (every call is one statistics domain)."""
        return any(self._normalised[:levels])

    def sync_features(self, level, feat):
        """feat: [1, C, H, W] tensor (any device/dtype)."""
        key = self._tensor_key([feat])
        if self._feat_key[level] == key:
            return
        if feat.dim() != 4 or feat.shape[0] != 1:
            raise ValueError("expected one [1, C, H, W] feature map, got %s" % (tuple(feat.shape),))
        f = feat.detach().to(self.device, torch.float32).contiguous()
        _lib.check(self.lib.pifu_set_features(self.h, level, ctypes.c_void_p(f.data_ptr()), f.shape[1],
                                              f.shape[2], f.shape[3], _stream(self.device_index)))
        self._keep[("feat", level)] = f        # keep alive until the copy on the stream is done
        self._hold[("feat", level)] = feat     # the keyed tensor itself (f may be a converted copy)
        self._feat_key[level] = key

    def set_options(self, perspective, load_size, z_size):
        key = (bool(perspective), int(load_size) // 2, float(z_size))
        if key != self._opt_key:
            _lib.check(self.lib.pifu_set_options(self.h, int(key[0]), float(key[1]), float(key[2])))
            self._opt_key = key

    def set_gemm_impl(self, impl):
        _lib.check(self.lib.pifu_set_gemm_impl(self.h, int(impl)))

    PRECISION = {"fast": 0, "split": 1, "hybrid": 2}

    def set_precision(self, mode="fast", terms=3, band=(0.02, 0.98)):
        """Arithmetic of the per-point MLP (`pifu_set_precision`): 'fast' (one fp16 image per tensor-core operand),
        'split' (fp16 + fp16 residual for features, activations and weights: fp32-level error, 3 tensor-core passes),
        'hybrid' (fast everywhere, points whose occupancy lies inside `band` again in split precision)."""
        m = self.PRECISION[mode] if isinstance(mode, str) else int(mode)
        _lib.check(self.lib.pifu_set_precision(self.h, m, int(terms), float(band[0]), float(band[1])))
        self.precision = m

    def refined_points(self):
        return int(self.lib.pifu_refined_points(self.h))

    def set_chunk_tiles(self, tiles):
        _lib.check(self.lib.pifu_set_chunk_tiles(self.h, int(tiles)))

    def profile(self, on):
        _lib.check(self.lib.pifu_profile_enable(self.h, int(bool(on))))

    def profile_read(self):
        """-> (launches, total_ms, total_flops) of the layer-kernel launches since profile(True)."""
        n, ms, fl = ctypes.c_longlong(), ctypes.c_double(), ctypes.c_double()
        _lib.check(self.lib.pifu_profile_read(self.h, ctypes.byref(n), ctypes.byref(ms), ctypes.byref(fl)))
        return n.value, ms.value, fl.value

    def profile_read_kind(self, kind):
        """-> (launches, total_ms, total_flops) of one kernel kind (0 layer kernel, 1 chain kernel on lattice
        columns, 2 chain kernel on id lists)."""
        n, ms, fl = ctypes.c_longlong(), ctypes.c_double(), ctypes.c_double()
        _lib.check(self.lib.pifu_profile_read_kind(self.h, int(kind), ctypes.byref(n), ctypes.byref(ms), ctypes.byref(fl)))
        return n.value, ms.value, fl.value

    def set_chain(self, enabled):
        """0 / False: per-layer kernels only; 1 / True: chain kernel for lattice columns (eval_grid) and for sorted
        id lists (octree frontiers, run-list form); 2: lattice-column form only."""
        _lib.check(self.lib.pifu_set_chain(self.h, int(enabled)))

    def chain_ready(self):
        return bool(self.lib.pifu_chain_ready(self.h))

    def launch_count(self):
        return int(self.lib.pifu_launch_count(self.h))

    # ------------------------------------------------------------------ compute
    def query(self, levels, points, calib_local, calib_global, want_low=False, want_phi=0, no_mask=False, precise=False):
        """points [3, n] fp32 on this device.  Returns (pred [n], low [n] | None, phi [C, n] | None)."""
        assert points.dim() == 2 and points.shape[0] == 3
        pts = points.detach().to(self.device, torch.float32)
        if pts.stride(1) != 1:
            pts = pts.contiguous()
        n = pts.shape[1]
        pred = torch.empty(n, device=self.device, dtype=torch.float32)
        low = torch.empty(n, device=self.device, dtype=torch.float32) if want_low else None
        phi = torch.empty(want_phi, n, device=self.device, dtype=torch.float32) if want_phi else None
        if n == 0:                     # e.g. gen_mesh's last colour chunk `left:-1` (reconstruction.py:64-68)
            return pred, low, phi
        cl, _ = _calib16(calib_local)
        cg, _ = _calib16(calib_global)
        _lib.check(self.lib.pifu_query(
            self.h, levels, (1 if no_mask else 0) | (2 if precise else 0), ctypes.c_void_p(pts.data_ptr()), pts.stride(0), n, cl, cg,
            ctypes.c_void_p(pred.data_ptr()),
            ctypes.c_void_p(low.data_ptr()) if low is not None else None,
            ctypes.c_void_p(phi.data_ptr()) if phi is not None else None,
            _stream(self.device_index)))
        return pred, low, phi

    @staticmethod
    def calib_pair(calib):
        """(float[16], double[16] inverse) exactly as `mesh_util.py:61-62` derives them."""
        c16, full = _calib16(calib)
        inv = np.linalg.inv(full.numpy())          # float32 in, like calib_tensor[0].cpu().numpy()
        inv64 = np.ascontiguousarray(inv, dtype=np.float64)
        return c16, (ctypes.c_double * 16)(*inv64.reshape(-1).tolist()), full, inv64

    def eval_grid(self, levels, res, calib, id_begin=0, id_end=None, out=None):
        R0, R1, R2 = (res, res, res) if np.isscalar(res) else res
        total = R0 * R1 * R2
        id_end = total if id_end is None else id_end
        if out is None:
            out = torch.empty(id_end - id_begin, device=self.device, dtype=torch.float32)
        c16, inv16, _, _ = self.calib_pair(calib)
        _lib.check(self.lib.pifu_eval_grid(self.h, levels, R0, R1, R2, id_begin, id_end, c16, inv16,
                                           ctypes.c_void_p(out.data_ptr()), _stream(self.device_index)))
        return out

    def eval_grid_host(self, levels, res, calib, out_host, id_begin=0, id_end=None, launches_per_piece=4):
        """eval_grid with the field delivered to HOST memory (`out_host`: pinned float32 tensor of id_end - id_begin
        entries; what `mesh_util.py:74` does per chunk with `.cpu()`).  The id range is evaluated in pieces of whole kernel
        launches (the lattice form cuts a call into launches of 32 x SM-count columns, api.cu) and a piece travels on a
        copy stream while the next one is computed, so only the last piece's transfer is exposed.  The calling stream
        waits for the copies: work queued after this call sees `out_host` complete (a host read still needs a sync)."""
        R0, R1, R2 = (res, res, res) if np.isscalar(res) else res
        id_end = R0 * R1 * R2 if id_end is None else id_end
        n = id_end - id_begin
        if out_host.numel() != n or out_host.dtype != torch.float32 or out_host.device.type != "cpu":
            raise ValueError("out_host must be a float32 host tensor of %d entries" % n)
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        piece = 32 * sms * R2 * max(1, int(launches_per_piece))
        dev = torch.empty(n, device=self.device, dtype=torch.float32)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs, main = self._copy_stream, torch.cuda.current_stream(self.device)
        flat = out_host.view(-1)
        for a in range(0, n, piece):
            b = min(a + piece, n)
            self.eval_grid(levels, res, calib, id_begin + a, id_begin + b, out=dev[a:b])
            done = torch.cuda.Event()
            done.record(main)
            cs.wait_event(done)
            with torch.cuda.stream(cs):
                flat[a:b].copy_(dev[a:b], non_blocking=True)
        main.wait_stream(cs)
        return out_host

    def eval_lattice_ids(self, levels, res, ids, calib):
        R0, R1, R2 = (res, res, res) if np.isscalar(res) else res
        ids = ids.to(self.device, torch.int64).contiguous()
        out = torch.empty(ids.numel(), device=self.device, dtype=torch.float32)
        c16, inv16, _, _ = self.calib_pair(calib)
        _lib.check(self.lib.pifu_eval_lattice_ids(self.h, levels, R0, R1, R2, ctypes.c_void_p(ids.data_ptr()),
                                                  ids.numel(), c16, inv16, ctypes.c_void_p(out.data_ptr()),
                                                  _stream(self.device_index)))
        return out

    def debug_gemm(self, X, W, b, leaky=True):
        X = X.to(self.device, torch.float32).contiguous()
        W = W.to(self.device, torch.float32).contiguous()
        b = b.to(self.device, torch.float32).contiguous()
        M, K = X.shape
        N = W.shape[0]
        Y = torch.empty(N, M, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.pifu_debug_gemm(self.h, ctypes.c_void_p(X.data_ptr()), ctypes.c_void_p(W.data_ptr()),
                                            ctypes.c_void_p(b.data_ptr()), M, K, N, int(leaky),
                                            ctypes.c_void_p(Y.data_ptr()), _stream(self.device_index)))
        return Y

    # ------------------------------------------------------------------ octree
    def eval_grid_octree(self, levels, res, calib, init_resolution=64, threshold=0.05,
                         want64=True, want32=False):
        """One-call device octree.  Returns (sdf64 | None, sdf32 | None, evaluated_per_level)."""
        R0, R1, R2 = (res, res, res) if np.isscalar(res) else res
        sdf64 = torch.empty((R0, R1, R2), device=self.device, dtype=torch.float64) if want64 else None
        sdf32 = torch.empty((R0, R1, R2), device=self.device, dtype=torch.float32) if want32 else None
        c16, inv16, _, _ = self.calib_pair(calib)
        stats = (ctypes.c_longlong * 16)()
        _lib.check(self.lib.pifu_eval_grid_octree(
            self.h, levels, R0, R1, R2, int(init_resolution), float(threshold), c16, inv16,
            ctypes.c_void_p(sdf64.data_ptr()) if want64 else None,
            ctypes.c_void_p(sdf32.data_ptr()) if want32 else None, stats, 16, _stream(self.device_index)))
        return sdf64, sdf32, [int(v) for v in stats if v >= 0]

    def octree_begin(self, res, init_resolution=64, threshold=0.05):
        R0, R1, R2 = (res, res, res) if np.isscalar(res) else res
        self._oct_res = (R0, R1, R2)
        _lib.check(self.lib.pifu_octree_begin(self.h, R0, R1, R2, int(init_resolution), float(threshold),
                                              _stream(self.device_index)))

    def octree_begin_slab(self, res, init_resolution, threshold, plane_begin, plane_end, own_begin, own_end):
        """Slab form (multi-GPU): bookkeeping of planes [plane_begin, plane_end), frontier of [own_begin, own_end)."""
        R0, R1, R2 = (res, res, res) if np.isscalar(res) else res
        self._oct_res = (plane_end - plane_begin, R1, R2)
        _lib.check(self.lib.pifu_octree_begin_slab(self.h, R0, R1, R2, int(init_resolution), float(threshold),
                                                   int(plane_begin), int(plane_end), int(own_begin), int(own_end),
                                                   _stream(self.device_index)))

    def octree_set_frontier_planes(self, plane_begin, plane_end):
        _lib.check(self.lib.pifu_octree_set_frontier_planes(self.h, int(plane_begin), int(plane_end)))

    def octree_commit_pairs(self, ids, vals):
        """(lattice id, value) pairs of all ranks' frontiers; those outside this rank's planes are ignored."""
        ids = ids.to(self.device, torch.int64).contiguous()
        vals = vals.to(self.device, torch.float32).contiguous()
        _lib.check(self.lib.pifu_octree_commit_pairs(self.h, ctypes.c_void_p(ids.data_ptr()) if ids.numel() else None,
                                                     ctypes.c_void_p(vals.data_ptr()) if vals.numel() else None,
                                                     ids.numel(), _stream(self.device_index)))

    def octree_field32(self):
        """-> (float32 [planes, R1, R2] view of the library's field, global index of its first plane)."""
        ptr, pb, n = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        _lib.check(self.lib.pifu_octree_field32(self.h, ctypes.byref(ptr), ctypes.byref(pb), ctypes.byref(n)))
        _, R1, R2 = self._oct_res
        t = torch.as_tensor(_DevView(ptr.value, n.value * R1 * R2, "<f4"), device=self.device)
        return t.view(n.value, R1, R2), pb.value

    def octree_frontier(self):
        """-> (step, ids) with ids a device int64 view of this level's lattice ids; step 0 = done."""
        n = ctypes.c_longlong()
        ptr = ctypes.c_void_p()
        step = ctypes.c_int()
        _lib.check(self.lib.pifu_octree_frontier(self.h, ctypes.byref(n), ctypes.byref(ptr), ctypes.byref(step),
                                                 _stream(self.device_index)))
        if step.value == 0 or n.value == 0:
            return step.value, torch.empty(0, device=self.device, dtype=torch.int64)
        # copy out: the library reuses its id buffer on the next level
        ids = torch.as_tensor(_DevView(ptr.value, n.value, "<i8"), device=self.device).clone()
        return step.value, ids

    def octree_commit(self, vals):
        """Values of the whole frontier, in frontier order; float64 tensors are stored as they are
        (generic callables), anything else as float32 (what `get_preds` returns)."""
        if vals.dtype == torch.float64:
            vals = vals.to(self.device).contiguous()
            fn = self.lib.pifu_octree_commit64
        else:
            vals = vals.to(self.device, torch.float32).contiguous()
            fn = self.lib.pifu_octree_commit
        _lib.check(fn(self.h, ctypes.c_void_p(vals.data_ptr()) if vals.numel() else None, _stream(self.device_index)))

    def octree_export(self, want64=True, want32=False):
        R0, R1, R2 = self._oct_res
        sdf64 = torch.empty((R0, R1, R2), device=self.device, dtype=torch.float64) if want64 else None
        sdf32 = torch.empty((R0, R1, R2), device=self.device, dtype=torch.float32) if want32 else None
        _lib.check(self.lib.pifu_octree_export(self.h, ctypes.c_void_p(sdf64.data_ptr()) if want64 else None,
                                               ctypes.c_void_p(sdf32.data_ptr()) if want32 else None,
                                               _stream(self.device_index)))
        return sdf64, sdf32

    # ------------------------------------------------------------------ marching cubes
    def _mc_extract(self, f, level, i_global0, global_n0, cell_layers, ghost, want_normals, key):
        """count + emit without a host synchronisation in between (`pifu_mc_extract`): the outputs are allocated
        from the sizes of the previous extraction of the same kind (+ 50 %); the counts come back with one read,
        and only an overflow repeats the call.  -> (verts, faces, normals, values, ghost_verts), exact-size views."""
        hint = self._mc_hint.get(key)
        if hint is None:
            hint = (64 * f.shape[1] * f.shape[2] // 8 + 4096, 128 * f.shape[1] * f.shape[2] // 8 + 8192)
        cap_v, cap_f = int(hint[0] * 3 // 2) + 1024, int(hint[1] * 3 // 2) + 2048
        while True:
            verts = torch.empty((cap_v, 3), device=self.device, dtype=torch.float64)
            faces = torch.empty((cap_f, 3), device=self.device, dtype=torch.int32)
            normals = torch.empty((cap_v, 3), device=self.device, dtype=torch.float32) if want_normals else None
            values = torch.empty((cap_v,), device=self.device, dtype=torch.float32) if want_normals else None
            counts = torch.empty(3, device=self.device, dtype=torch.int64)
            _lib.check(self.lib.pifu_mc_extract(
                self.h, ctypes.c_void_p(f.data_ptr()), f.shape[0], f.shape[1], f.shape[2], float(level),
                int(i_global0), int(global_n0), int(cell_layers), 1 if ghost else 0,
                ctypes.c_void_p(verts.data_ptr()), ctypes.c_void_p(faces.data_ptr()),
                ctypes.c_void_p(normals.data_ptr()) if want_normals else None,
                ctypes.c_void_p(values.data_ptr()) if want_normals else None,
                cap_v, cap_f, ctypes.c_void_p(counts.data_ptr()), _stream(self.device_index)))
            nv, nf, ng = (int(x) for x in counts.tolist())
            self._mc_hint[key] = (nv, nf)
            if nv <= cap_v and nf <= cap_f:
                break
            cap_v, cap_f = max(cap_v, nv), max(cap_f, nf)
        self._keep["mc_field"] = f
        return (verts[:nv], faces[:nf], normals[:nv] if want_normals else None,
                values[:nv] if want_normals else None, ng)

    def marching_cubes_slab_async(self, field, level, i_global0, global_n0, cell_layers, ghost, cap=None):
        """Slab extraction without any host synchronisation: -> (verts, faces, normals, values, counts) with the outputs at
        capacity `cap` = (verts, faces) (default: from the previous extraction of this slab) and counts a device int64 [3]
        = vertices, faces, ghost vertices.  The caller reads the counts (e.g. together with the other ranks') and calls
        again with a larger `cap` if one was exceeded."""
        f = field.to(self.device, torch.float32).contiguous()
        key = ("slab", int(i_global0), int(cell_layers)) + tuple(f.shape)
        if cap is None:
            hint = self._mc_hint.get(key) or (64 * f.shape[1] * f.shape[2] // 8 + 4096, 128 * f.shape[1] * f.shape[2] // 8 + 8192)
            cap = (int(hint[0] * 3 // 2) + 1024, int(hint[1] * 3 // 2) + 2048)
        cap_v, cap_f = int(cap[0]), int(cap[1])
        verts = torch.empty((cap_v, 3), device=self.device, dtype=torch.float64)
        faces = torch.empty((cap_f, 3), device=self.device, dtype=torch.int32)
        normals = torch.empty((cap_v, 3), device=self.device, dtype=torch.float32)
        values = torch.empty((cap_v,), device=self.device, dtype=torch.float32)
        counts = torch.empty(3, device=self.device, dtype=torch.int64)
        _lib.check(self.lib.pifu_mc_extract(
            self.h, ctypes.c_void_p(f.data_ptr()), f.shape[0], f.shape[1], f.shape[2], float(level),
            int(i_global0), int(global_n0), int(cell_layers), 1 if ghost else 0,
            ctypes.c_void_p(verts.data_ptr()), ctypes.c_void_p(faces.data_ptr()), ctypes.c_void_p(normals.data_ptr()),
            ctypes.c_void_p(values.data_ptr()), cap_v, cap_f, ctypes.c_void_p(counts.data_ptr()), _stream(self.device_index)))
        self._keep["mc_field"] = f
        self._mc_last_key = key
        return verts, faces, normals, values, counts

    def marching_cubes_note_counts(self, nv, nf):
        """Remember the sizes of the last asynchronous slab extraction (its next capacity)."""
        self._mc_hint[self._mc_last_key] = (int(nv), int(nf))

    def marching_cubes(self, field, level, want_normals=True):
        """field: device float32 [n0, n1, n2].  -> (verts f64 [V,3], faces i32 [F,3], normals, values)
        on the device.  Raises ValueError like skimage when there is no surface at `level`."""
        f = field.to(self.device, torch.float32).contiguous()
        if f.dim() != 3:
            raise ValueError("Input volume should be a 3D array")
        verts, faces, normals, values, _ = self._mc_extract(f, level, 0, f.shape[0], f.shape[0] - 1, False, want_normals,
                                                            ("whole",) + tuple(f.shape))
        if verts.shape[0] == 0:
            raise ValueError("No surface found at the given iso value (or level outside the data range)")
        return verts, faces, normals, values

    def marching_cubes_slab(self, field, level, i_global0, global_n0, cell_layers, ghost, want_normals=True):
        """Slab form (multi-GPU): `field` = planes [i_global0, i_global0 + n0) of a global_n0-plane
        volume; the first `cell_layers` cell layers are processed, the first one only as a ghost when
        `ghost`.  -> (verts, faces, normals, values, ghost_verts): the caller drops the first
        ghost_verts vertices and renumbers faces by (first own global vertex number - ghost_verts).
        An empty slab returns zero-length tensors (a neighbour may still hold the surface)."""
        f = field.to(self.device, torch.float32).contiguous()
        return self._mc_extract(f, level, i_global0, global_n0, cell_layers, ghost, want_normals,
                                ("slab", int(i_global0), int(cell_layers)) + tuple(f.shape))


    # ------------------------------------------------------------------ colours, cleaning
    def sample_image(self, image, points, calib, perspective=False):
        """image [C, H, W] (or [1, C, H, W]), points [3, n], calib [4, 4] -> [C, n] bilinear samples at the projected
        points (`pifu_sample_image`): projection + `index()` of `reconstruction.py:110-116` in one kernel."""
        img = image.detach().to(self.device, torch.float32)
        img = img[0] if img.dim() == 4 else img
        img = img.contiguous()
        pts = points.detach().to(self.device, torch.float32)
        if pts.stride(1) != 1:
            pts = pts.contiguous()
        n = pts.shape[1]
        out = torch.empty((img.shape[0], n), device=self.device, dtype=torch.float32)
        c16, _ = _calib16(calib)
        _lib.check(self.lib.pifu_sample_image(ctypes.c_void_p(img.data_ptr()), img.shape[0], img.shape[1], img.shape[2],
                                              ctypes.c_void_p(pts.data_ptr()), pts.stride(0), n, c16, int(bool(perspective)),
                                              ctypes.c_void_p(out.data_ptr()), _stream(self.device_index)))
        return out

    def clean_mesh(self, verts, faces, colors=None, only_watertight=True):
        """Largest component by extent along axis 0 (`pifu_mesh_clean`; `reconstruction.py:325-344`).
        -> (verts [V', 3] f64, faces [F', 3] i32, colors [V', 3] f64 | None) on the device."""
        v = verts.to(self.device, torch.float64).contiguous()
        f = faces.to(self.device, torch.int32).contiguous()
        c = colors.to(self.device, torch.float64).contiguous() if colors is not None else None
        ov, of = torch.empty_like(v), torch.empty_like(f)
        oc = torch.empty_like(c) if c is not None else None
        counts = (ctypes.c_longlong * 2)()
        _lib.check(self.lib.pifu_mesh_clean(ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(c.data_ptr()) if c is not None else None,
                                            ctypes.c_void_p(f.data_ptr()), v.shape[0], f.shape[0], int(bool(only_watertight)),
                                            ctypes.c_void_p(ov.data_ptr()), ctypes.c_void_p(oc.data_ptr()) if oc is not None else None,
                                            ctypes.c_void_p(of.data_ptr()), counts, _stream(self.device_index)))
        nv, nf = int(counts[0]), int(counts[1])
        return ov[:nv], of[:nf], (oc[:nv] if oc is not None else None)


class _DevView:
    """Zero-copy view of library-owned device memory through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
