#!/usr/bin/env python
"""Benchmark of the reconstruction hot path (BASELINE.json: occupancy queries/s; configs[1] =
PIFuMRNet multi-level, dense 256^3 lattice, per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one dense lattice: in-kernel lattice generation -> calib
projection -> bilinear sampling of the coarse/fine feature maps -> coarse MLP trunk -> fine MLP ->
occupancy field in HBM.  N = 1: the 256^3 lattice of configs[1] (16 777 216 queries).  N > 1 (torchrun, one
rank per GPU): ONE 512^3 lattice (configs[2]'s volume, 134 217 728 queries) slab-sharded along axis 0, no
data-path collective; the slabs are gathered to rank 0 with NCCL inside the timed step (strong scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES = 256
METRIC = "occupancy_queries_per_s"
UNIT = "queries/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d["bf16_tflops_sustained"]), hbm=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu
    capture of the same workload shape (profiles/ncu_traffic.json); None when there is none."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p)).get(kernel)
        return None if not d else {"bytes_per_launch": d["dram_bytes_read"] + d["dram_bytes_write"],
                                   "queries_per_launch": d["queries_per_launch"], "source": d["source"]}
    except (OSError, ValueError, KeyError):
        return None


def ncu_mc_traffic():
    """DRAM bytes of the marching-cubes launches of one 512^3 extraction on the bench's octree field (committed ncu launch
    list, profiles/ncu_traffic.json); None when there is none."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("marching_cubes")
        return None if not d else {"bytes_per_extraction": d["dram_bytes_per_extraction"], "verts": d["verts"],
                                   "faces": d["faces"], "source": d["source"]}
    except (OSError, ValueError, KeyError):
        return None


def build_problem():
    """Seeded random-init two-level net + band-limited feature maps (pifu_b200.synthetic)."""
    from pifu_b200 import synthetic as syn
    prob = syn.make_problem(bias_std=0.0)
    return prob, syn.default_calib()


def calibrate_from_pilot(prob, pilot_preds, saturated=False):
    """the raw random-init field is a Gaussian of width 4.5e-4 around 0.5 (no surface, every sign a
    coin flip).  Only the last fine conv is rescaled, from the un-calibrated net's predictions on seeded pilot
    points, so that 2 % of the cube exceeds 0.5 with O(1) logits (the parity-gate field); `saturated` adds the x8
    gain of the octree / marching-cubes field.  The same tensors then go to every implementation."""
    from pifu_b200 import synthetic as syn
    syn.calibrate_last_layer(prob["fine"], 3, pilot_preds)
    if saturated:
        syn.saturate(prob["fine"], 3)
    return prob


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms from before the warm-up; only the
    samples whose nvidia-smi timestamp falls inside the timed region are reported."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark_begin(self):
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        self.t1 = datetime.datetime.now()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                if self.t0 and self.t1 and not (self.t0 <= ts <= self.t1):
                    continue
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for n, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def sample_ids(res, n_points):
    """The bounded sample of the res^3 lattice the CPU legs evaluate: every (res^3 // n_points)-th lattice id."""
    stride = max(1, res ** 3 // n_points)
    return np.arange(0, res ** 3, stride)[:n_points]


def oracle_fine_state(prob):
    from pifu_b200 import config
    from oracle import pifu_oracle as orc
    oc, of = config.coarse_opt(), config.fine_opt()
    coarse = orc.CoarseState(prob["coarse"], prob["feat_coarse"], oc)
    return orc.FineState(prob["fine"], prob["feat_fine"], of, coarse)


def cpu_port_queries_per_s(prob, calib, n_points, steps=1, warmup=0, res=RES, keep=False):
    """Times the CPU oracle (torch-CPU port of PIFuMRNet.query, same library kernels the
    reference runs) on a bounded sample of the same lattice: every (res^3 // n_points)-th point.
    With `keep` the occupancies of the sample are returned as well (the parity block compares the
    GPU field with them)."""
    from oracle import pifu_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    fine = oracle_fine_state(prob)
    ids = sample_ids(res, n_points)
    k, j, i = ids % res, (ids // res) % res, ids // (res * res)
    pts = np.stack([-1 + 2.0 * i / res, -(-1 + 2.0 * j / res), -1 + 2.0 * k / res]).astype(np.float32)
    pts = torch.from_numpy(pts)[None]
    chunk = 100000                      # gen_mesh_imgColor's num_samples (reconstruction.py:108)
    times, vals = [], None
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out = [orc.query_fine(fine, pts[:, :, s:s + chunk], calib)[0] for s in range(0, pts.shape[2], chunk)]
            if it >= warmup:
                times.append(time.perf_counter() - t0)
            if keep:
                vals = torch.cat(out, 2).numpy().ravel()
    sec = sum(times) / len(times)
    res_t = (pts.shape[2] / sec, torch.get_num_threads(), pts.shape[2], sec)
    return res_t + (vals,) if keep else res_t


def parity_block(out, ref, what):
    """GPU field vs the CPU oracle on the same lattice points (north_star: <= 1e-3 absolute, >= 99.99 % sign
    agreement at the 0.5 iso-level, identical in-bounds masks)."""
    out, ref = np.asarray(out, dtype=np.float32).ravel(), np.asarray(ref, dtype=np.float32).ravel()
    return {"n": int(ref.size), "max_abs_err": float(np.abs(out - ref).max()),
            "sign_agreement": float(((out > 0.5) == (ref > 0.5)).mean()),
            "mask_equal": bool(np.array_equal(out == 0, ref == 0)),
            "occupied_fraction": float((ref > 0.5).mean()), "what": what}


def build_nets(dev, saturated, with_prob=False):
    """Two-level net whose random-init field has an iso-surface: only the last
    fine conv is rescaled, from the un-calibrated net's own predictions on seeded pilot points
    (fast arithmetic on the GPU; the raw logits are ~1e-3, so their fp16 noise is ~1e-6 of the
    calibrated spread), so ~2 % of the cube is occupied; `saturated` adds the x8 gain that makes
    the sigmoid saturate away from the surface (octree / marching-cubes field)."""
    from pifu_b200 import PIFuMRNet, PIFuNetwNML, config, synthetic as syn
    prob = syn.make_problem(bias_std=0.0)
    calib = syn.default_calib()
    netG = PIFuNetwNML(config.coarse_opt(), "orthogonal")
    netMR = PIFuMRNet(config.fine_opt(), netG, "orthogonal")
    netG.mlp.load_state_dict(prob["coarse"])
    netMR.mlp.load_state_dict(prob["fine"])
    netMR.to(dev).eval()
    netG.im_feat_list = [prob["feat_coarse"].to(dev)]
    netMR.im_feat_list = [prob["feat_fine"].to(dev)]
    pilot = syn.random_points(20000, syn.SEED_PILOT, -1.0, 1.0)
    eng = netMR._engine_for(torch.zeros(1, device=dev))
    eng.set_precision("split")                  # the pilot in fp32-level arithmetic: the calibration is then the same
    netMR.query(pilot.to(dev), calib.to(dev))   # tensor the oracle would derive (a scale and a bias of one layer)
    eng.set_precision("fast")
    p = netMR.get_preds().float().cpu().numpy()
    calibrate_from_pilot(prob, p, saturated)
    netMR.mlp.load_state_dict(prob["fine"])
    netMR.to(dev).eval()
    eng = netMR._engine_for(torch.zeros(1, device=dev))
    eng.sync_features(0, netG.im_feat_list[-1])
    eng.sync_features(1, netMR.im_feat_list[-1])
    if with_prob:
        return netG, netMR, eng, calib, prob
    return netG, netMR, eng, calib


def build_mesh_problem(dev):
    return build_nets(dev, saturated=True)


def group_norm_leg(dev, res, chunks=32, num_samples=100000):
    """The reference's DEFAULT mlp_norm ('group', options.py:95): GroupNorm(32) between every Conv1d and its
    leaky_relu, statistics over the points of one query() call - so the lattice is cut into the reference's own
    calls (num_samples = 100000, reconstruction.py:108,160), each one statistics domain through the per-layer
    kernels + norm.cu.  Reported beside the headline, which is quoted on mlp_norm 'none'."""
    from pifu_b200 import PIFuMRNet, PIFuNetwNML, config, synthetic as syn
    prob = syn.make_problem(bias_std=0.0)
    netG = PIFuNetwNML(config.coarse_opt(mlp_norm="group"), "orthogonal")
    netMR = PIFuMRNet(config.fine_opt(mlp_norm="group"), netG, "orthogonal")
    netG.mlp.load_state_dict(prob["coarse"], strict=False)          # GroupNorm affine stays at its init (1, 0)
    netMR.mlp.load_state_dict(prob["fine"], strict=False)
    netMR.to(dev).eval()
    netG.im_feat_list = [prob["feat_coarse"].to(dev)]
    netMR.im_feat_list = [prob["feat_fine"].to(dev)]
    eng = netMR._engine_for(torch.zeros(1, device=dev))
    eng.sync_features(0, netG.im_feat_list[-1])
    eng.sync_features(1, netMR.im_feat_list[-1])
    calib = syn.default_calib()
    out = torch.empty(num_samples, device=dev, dtype=torch.float32)
    first = (res ** 3) // 3                     # a run of calls in the middle of the lattice

    def run():
        for c in range(chunks):
            b = first + c * num_samples
            eng.eval_grid(2, res, calib[0], id_begin=b, id_end=b + num_samples, out=out)

    run()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    run()
    ev[1].record()
    torch.cuda.synchronize(dev)
    ms = ev[0].elapsed_time(ev[1])
    return {"mlp_norm": "group", "queries_per_s": chunks * num_samples / (ms * 1e-3), "calls": chunks,
            "points_per_call": num_samples, "ms": ms,
            "note": "per-layer tcgen05 kernels + call-wide GroupNorm statistics (norm.cu); not shardable, not chained"}


def mesh_cpu_baseline(prob, calib, res):
    """BASELINE's second figure on the host: the reference's CPU reconstruction path (`mesh_util.py:40-96` with
    use_octree=True) restated by the oracle - create_grid + calib pre-transform, eval_grid_octree with the
    torch-CPU query as eval_func (num_samples = 100000), marching cubes (oracle/mc_ref.c; scikit-image is absent).
    Returns the timings and the field / mesh sizes the GPU results are compared with."""
    from oracle import mc_oracle
    from oracle import pifu_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    fine = oracle_fine_state(prob)
    t0 = time.perf_counter()
    coords, _, _ = orc.lattice_coords(res, calib)
    t1 = time.perf_counter()
    stats = []
    ef = orc.make_eval_func(lambda p, c: orc.query_fine(fine, p, c), calib)
    with torch.no_grad():
        sdf = orc.eval_grid_octree(coords, ef, num_samples=100000, stats=stats)
    del coords
    t2 = time.perf_counter()
    verts, faces, _, _, _ = mc_oracle.marching_cubes(sdf, 0.5)
    t3 = time.perf_counter()
    return {"sdf": sdf, "evaluated_per_level": [n for _, n in stats], "verts": len(verts), "faces": len(faces),
            "create_grid_ms": (t1 - t0) * 1e3, "octree_ms": (t2 - t1) * 1e3, "mc_ms": (t3 - t2) * 1e3,
            "total_ms": (t3 - t0) * 1e3, "cores": torch.get_num_threads()}


def mesh_latency(netMR, eng, calib, dev, res=512, reps=3, cpu=None):
    """BASELINE.json's second figure: end-to-end latency of mesh_util.reconstruction at res^3 on
    this GPU (features resident; field -> marching cubes -> mesh on the host), octree and dense,
    with the phases timed by CUDA events and the marching-cubes HBM roofline.  `cpu`: the host's
    result of the same reconstruction (mesh_cpu_baseline) the octree fields are compared with."""
    from pifu_b200 import mesh_util
    peaks = load_peaks()
    cal = calib.to(dev)
    out = {"resolution": res}
    # the same reconstruction in the hybrid arithmetic (parity gates met as stated on this saturated field), and with
    # the band narrowed to the occupancies whose SIGN the fast arithmetic could get wrong (|p - 0.5| < 0.1)
    for key, band in (("octree_hybrid_narrow_band", (0.4, 0.6)), ("octree_hybrid", (0.02, 0.98))):
        eng.set_precision("hybrid", band=band)
        try:
            hy = {"band": list(band)}
            best = None
            r0 = eng.refined_points()
            for _ in range(reps + 1):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                mesh_h = mesh_util.reconstruction(netMR, dev, cal, res, None, None, thresh=0.5, use_octree=True, num_samples=5000)
                torch.cuda.synchronize(dev)
                dt = (time.perf_counter() - t0) * 1e3
                best = dt if best is None else min(best, dt)
            hy["latency_ms"] = best
            hy["refined_points_per_reconstruction"] = (eng.refined_points() - r0) // (reps + 1)
            hy["verts"], hy["faces"] = (len(mesh_h[0]), len(mesh_h[1])) if mesh_h != -1 else (0, 0)
            if cpu is not None:
                st = []
                f = mesh_util.eval_field_device(netMR, dev, cal, res, True, stats=st)
                hy["vs_cpu"] = field_vs_cpu(f, st, cpu)
                del f
            out[key] = hy
        finally:
            eng.set_precision("fast")
    for mode in ("octree", "dense"):
        best = None
        for _ in range(reps + 1):                      # first pass warms allocations
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            mesh = mesh_util.reconstruction(netMR, dev, cal, res, None, None, thresh=0.5,
                                            use_octree=(mode == "octree"), num_samples=5000)
            torch.cuda.synchronize(dev)
            dt = (time.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        stats = []
        ev[0].record()
        field = mesh_util.eval_field_device(netMR, dev, cal, res, mode == "octree", stats=stats)
        ev[1].record()
        l0 = eng.launch_count()
        verts, faces, normals, values = eng.marching_cubes(field, 0.5)
        ev[2].record()
        host = mesh_util._to_host(verts, faces, normals, values)       # pinned, as reconstruction() does
        ev[3].record()
        torch.cuda.synchronize(dev)
        mc_ms = ev[1].elapsed_time(ev[2])
        nv, nf = int(verts.shape[0]), int(faces.shape[0])
        mc_bytes = 4.0 * res ** 3 + nv * (24 + 12 + 4) + nf * 12
        d = {"latency_ms": best, "field_ms": ev[0].elapsed_time(ev[1]), "mc_ms": mc_ms,
             "mesh_d2h_ms": ev[2].elapsed_time(ev[3]), "verts": nv, "faces": nf,
             "mc_launches": eng.launch_count() - l0,
             "mc_roofline": {"bound": "hbm", "achieved": mc_bytes / (mc_ms * 1e-3) / 1e9, "peak": peaks["hbm"],
                             "unit": "GB/s", "frac": mc_bytes / (mc_ms * 1e-3) / 1e9 / peaks["hbm"],
                             "algorithmic_bytes": mc_bytes,
                             "traffic": (ncu_mc_traffic() or {}).get("bytes_per_extraction") if (mode == "octree" and res == 512) else None,
                             "traffic_detail": ncu_mc_traffic() if (mode == "octree" and res == 512) else None,
                             "note": "4 B/voxel field + 40 B/vertex (f64 position, f32 normal, value) + 12 B/face; "
                                     "count + emit, includes the host sync that sizes the output"}}
        if mode == "octree" and mesh != -1:
            # the rest of the reference's gen_mesh (`reconstruction.py:58-72`): normals as
            # colours, 4 finite-difference queries per vertex in chunks of 50 000, then the OBJ text file
            import tempfile
            mv, mf = mesh[0], mesh[1]

            def tail():
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                verts_tensor = torch.from_numpy(mv.T).unsqueeze(0).to(device=dev).float()
                color = np.zeros(mv.shape)
                interval = 50000
                for i in range(len(color) // interval + 1):
                    left = i * interval
                    right = -1 if i == len(color) // interval else (i + 1) * interval
                    netMR.calc_normal(verts_tensor[:, None, :, left:right], cal[:, None], cal)
                    color[left:right] = (netMR.nmls.detach().cpu().numpy()[0] * 0.5 + 0.5).T
                t1 = time.perf_counter()
                with tempfile.TemporaryDirectory() as tmp:
                    path = os.path.join(tmp, "mesh.obj")
                    mesh_util.save_obj_mesh_with_color(path, mv, mf, color)
                    t2 = time.perf_counter()
                    nbytes = os.path.getsize(path)
                return t0, t1, t2, nbytes

            tail()                                       # first call: lazy initialisations of the general-points path
            t0, t1, t2, nbytes = tail()
            d["gen_mesh_tail"] = {"vertex_normals_ms": (t1 - t0) * 1e3, "queries": 4 * len(mv), "obj_write_ms": (t2 - t1) * 1e3,
                                  "obj_bytes": nbytes, "gen_mesh_total_ms": best + (t2 - t0) * 1e3,
                                  "normals_arithmetic": "split precision (finite differences at delta = 0.001)"}
            # colours sampled from the 1024^2 image (gen_mesh_imgColor, reconstruction.py:110-116) and
            # meshcleaning (reconstruction.py:325-344), host arrays in / host arrays out
            img = torch.rand(1, 3, 1024, 1024, device=dev) * 2 - 1
            for _ in range(2):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                colors = mesh_util.vertex_colors_from_image(netMR, img, mv, cal)
                t1 = time.perf_counter()
                try:
                    cv, cf, _ = mesh_util.clean_mesh(mv, mf, colors, device=dev)
                    kept = [len(cv), len(cf)]
                except Exception as e:  # noqa: BLE001
                    kept = "none (%s)" % (str(e)[:80],)
                t2 = time.perf_counter()
            d["postprocess"] = {"vertex_colors_ms": (t1 - t0) * 1e3, "meshcleaning_ms": (t2 - t1) * 1e3, "kept_verts_faces": kept}
            # the reference's own form: meshcleaning(obj_path), OBJ file in, OBJ file out (`reconstruction.py:325-344`)
            try:
                import contextlib
                import io
                with tempfile.TemporaryDirectory() as tmp:
                    path = os.path.join(tmp, "mesh.obj")
                    mesh_util.save_obj_mesh_with_color(path, mv, mf, colors)
                    t0 = time.perf_counter()
                    lv, lf, lc = mesh_util.load_obj_mesh_with_color(path)
                    t1 = time.perf_counter()
                    with contextlib.redirect_stdout(io.StringIO()):
                        mesh_util.meshcleaning(path, device=dev)
                    t2 = time.perf_counter()
                d["postprocess"]["obj_read_ms"] = (t1 - t0) * 1e3
                d["postprocess"]["meshcleaning_file_ms"] = (t2 - t1) * 1e3
            except Exception as e:  # noqa: BLE001
                d["postprocess"]["meshcleaning_file_ms"] = "failed (%s)" % (str(e)[:80],)
        if mode == "octree" and cpu is not None:
            d["vs_cpu"] = field_vs_cpu(field, stats, cpu)
        if mode == "octree":
            d["evaluated_per_level"] = stats
            d["evaluated_fraction"] = sum(stats) / float(res ** 3)
            # how the frontier points were evaluated: chain kernel in its run-list form (kind 2) and/or
            # per-layer kernels (kind 0, which also computes the per-run constants), per-launch CUDA events
            eng.profile(True)
            mesh_util.eval_field_device(netMR, dev, cal, res, True)
            n2, ms2, fl2 = eng.profile_read_kind(2)
            n0, ms0, fl0 = eng.profile_read_kind(0)
            eng.profile(False)
            d["frontier_mlp"] = {"runlist_chain_launches": n2, "runlist_chain_ms": ms2,
                                 "runlist_chain_algorithmic_tflops": (fl2 / (ms2 * 1e-3) / 1e12) if ms2 > 0 else None,
                                 "layer_kernel_launches": n0, "layer_kernel_ms": ms0}
        out[mode] = d
        del field, verts, faces, normals, values, host
    return out


def field_vs_cpu(field, stats, cpu):
    """GPU octree field (device float32 [R, R, R]) against the host's float64 field of the same reconstruction."""
    ref = torch.from_numpy(cpu["sdf"]).to(field.device)
    diff = (field.double() - ref).abs()
    agree = ((field > 0.5) == (ref > 0.5)).double().mean().item()
    over = (diff > 1e-3).double().mean().item()
    r = {"evaluated_per_level_gpu": list(stats), "evaluated_per_level_cpu": cpu["evaluated_per_level"],
         "evaluated_sets_equal_in_size": list(stats) == list(cpu["evaluated_per_level"]),
         "sign_agreement_all_voxels": agree, "max_abs_err": diff.max().item(), "fraction_over_1e-3": over,
         "voxels": int(ref.numel())}
    del ref, diff
    return r


def mesh_latency_sharded(netMR, calib, dev, res=512, reps=3, modes=("octree", "dense")):
    """The same figure at N > 1 (one rank per GPU): dense lattices are cut into slabs along axis 0, an octree
    level's frontier into equal shares; every rank extracts the iso-surface of its own slab after one halo
    exchange and the fragments are gathered on rank 0 (pifu_b200.dist).  Wall clock between barriers, best of
    `reps`; every rank returns the same dict."""
    import torch.distributed as dist
    from pifu_b200 import mesh_util
    cal = calib.to(dev)
    out = {"resolution": res, "ranks": dist.get_world_size()}
    for mode in modes:
        best, shape = None, None
        for _ in range(reps + 1):
            dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            mesh = mesh_util.reconstruction(netMR, dev, cal, res, None, None, thresh=0.5,
                                            use_octree=(mode == "octree"), num_samples=5000)
            torch.cuda.synchronize(dev)
            dist.barrier()
            dt = (time.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
            if mesh is not None and mesh != -1:
                shape = (len(mesh[0]), len(mesh[1]))
        t = torch.tensor([best], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[mode] = {"latency_ms": float(t.item()), "verts": shape[0] if shape else None, "faces": shape[1] if shape else None}
    return out


def encoder_leg(dev, frames=8, time_modes=True, world=1):
    """SURVEY §8(f) row 3 / BASELINE configs[3]: the per-image PyTorch encoders (coarse 4-stack hourglass on
    512^2 + fine 1-stack on 1024^2, 6-channel RGB-D input) timed per execution mode, and the whole per-frame
    flow - images H2D from pinned memory, filter_global, filter_local, 256^3 octree reconstruction, mesh on
    the host - in frames/s on this GPU (frames shard by frame across GPUs with no communication)."""
    from pifu_b200 import PIFuMRNet, PIFuNetwNML, config, mesh_util, synthetic as syn
    og = config.coarse_opt(use_front_normal=True)               # norm 'batch' (options.py:78), eval mode
    netG = PIFuNetwNML(og, "orthogonal")
    netG.netF = None                                            # RGB-D: depth rides in the normal channels
    netMR = PIFuMRNet(config.fine_opt(), netG, "orthogonal")
    prob = syn.make_problem(bias_std=0.0)
    netG.mlp.load_state_dict(prob["coarse"])
    netMR.mlp.load_state_dict(prob["fine"])
    syn.fill_state(netG.image_filter, 3)
    syn.fill_state(netMR.image_filter, 4)
    netMR.to(dev).eval()
    img512 = [syn.encoder_input((1, 6, 512, 512), 100 + f).pin_memory() for f in range(2)]
    img1024 = [syn.encoder_input((1, 1, 6, 1024, 1024), 200 + f).pin_memory() for f in range(2)]
    calib = syn.default_calib().to(dev)
    out = {"input": "6-channel 512^2 + 1024^2 per frame", "modes": {}}

    def time_filters(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        a, b = img512[0].to(dev), img1024[0].to(dev)
        for _ in range(2):
            netMR.filter_global(a); netMR.filter_local(b)
        ev[0].record()
        for _ in range(reps):
            netMR.filter_global(a); netMR.filter_local(b)
        ev[1].record()
        torch.cuda.synchronize(dev)
        return ev[0].elapsed_time(ev[1]) / reps

    for name, opts in () if not time_modes else (("eager_fp32_nchw", dict(channels_last=False, precision="fp32", graph=False, autotune=False)),
                       ("eager_tf32_nchw (stock PyTorch defaults)", dict(channels_last=False, precision="tf32", graph=False, autotune=False)),
                       ("fp32_nchw_graph", dict(channels_last=False, precision="fp32", graph=True)),
                       ("tf32_nchw_graph (default)", dict(channels_last=False, precision="tf32", graph=True)),
                       ("bf16_nchw_graph", dict(channels_last=False, precision="bf16", graph=True)),
                       ("tf32_channels_last_graph", dict(channels_last=True, precision="tf32", graph=True)),
                       ("bf16_channels_last_graph", dict(channels_last=True, precision="bf16", graph=True))):
        for net in (netG, netMR):
            net.reset_encoder_runners()
            net.encoder_opts = opts
        out["modes"][name] = {"filter_global_plus_local_ms": time_filters(5)}
    # per-frame flow with the default (tf32, NCHW, graph) encoders; field calibrated on the first frame
    for net in (netG, netMR):
        net.reset_encoder_runners()
        net.encoder_opts = None
        net.image_filter.to(memory_format=torch.contiguous_format)
    netMR.filter_global(img512[0].to(dev)); netMR.filter_local(img1024[0].to(dev))
    pilot = syn.random_points(20000, syn.SEED_PILOT, -1.0, 1.0)
    netMR.query(pilot.to(dev), calib)
    syn.calibrate_last_layer(prob["fine"], 3, netMR.get_preds().float().cpu().numpy())
    syn.saturate(prob["fine"], 3)
    netMR.mlp.load_state_dict(prob["fine"])
    netMR.to(dev).eval()

    solo = None                                     # N > 1: every rank reconstructs its own frames (a one-rank group)
    if world > 1:
        import torch.distributed as dist
        for r in range(world):
            g = dist.new_group([r])
            if r == dist.get_rank():
                solo = g

    def frame(f):
        netMR.filter_global(img512[f % 2].to(dev, non_blocking=True))
        netMR.filter_local(img1024[f % 2].to(dev, non_blocking=True))
        return mesh_util.reconstruction(netMR, dev, calib, 256, None, None, thresh=0.5, use_octree=True, num_samples=5000,
                                        group=solo)

    for f in range(2):
        mesh = frame(f)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for f in range(frames):
        mesh = frame(f)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    if world > 1:                                   # frames shard by frame, no communication: the box's rate is set by the slowest rank
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    out["frames_256_octree"] = {"frames": frames * world, "frames_per_s": frames * world / dt, "ms_per_frame": dt / frames * 1e3,
                                "ranks": world,
                                "mesh": None if mesh == -1 else [len(mesh[0]), len(mesh[1])],
                                "h2d_bytes_per_frame": (6 * 512 * 512 + 6 * 1024 * 1024) * 4}
    return out


def workload(world):
    """(lattice resolution, description) of the timed step: configs[1] on one GPU, one configs[2]-sized volume
    slab-sharded over N > 1 GPUs."""
    if world == 1:
        return RES, ("PIFuMRNet multi-level (coarse 257-1024-512-256 trunk + fine 272-512-256-128-1), dense %d^3 lattice = "
                     "%d queries on 1 GPU (BASELINE configs[1])" % (RES, RES ** 3))
    return 512, ("PIFuMRNet multi-level (coarse 257-1024-512-256 trunk + fine 272-512-256-128-1), ONE dense 512^3 lattice = "
                 "%d queries slab-sharded along axis 0 over %d GPUs (BASELINE configs[2]'s volume, dense)" % (512 ** 3, world))


def config_dict(world, res, what):
    """The `config` object, identical in both arms (--impl ours / reference)."""
    return {"workload": what, "queries_per_step": res ** 3, "mlp_norm": "none",
            "field": "random-init weights, band-limited synthetic feature maps, last fine conv calibrated to logit sigma 0.75 / "
                     "2 % occupied (SURVEY 7.3-2)",
            "l2": "GPU arm: 256 MiB flush write between timed iterations; CPU arm: not applicable"}


def run_reference(args):
    """--impl reference: the reference's own CPU path for this metric (oracle port: the
    reference cannot travel to the GPU box and its query path is library torch ops)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    res, what = workload(world)
    prob, calib = build_problem()
    # the same calibration rule as the GPU arm, from the oracle's own pilot predictions
    from oracle import pifu_oracle as orc
    from pifu_b200 import synthetic as syn
    with torch.no_grad():
        pilot = orc.query_fine(oracle_fine_state(prob), syn.random_points(20000, syn.SEED_PILOT, -1.0, 1.0), calib)[0].numpy()
    calibrate_from_pilot(prob, pilot)
    n = 200000
    qps, cores, npts, sec = cpu_port_queries_per_s(prob, calib, n, steps=args.steps, warmup=args.warmup, res=res)
    sample = "%d-point strided sub-lattice of the %d^3 lattice per step, chunks of 100000 (PIFuMRNet.query port, torch CPU fp32)" % (npts, res)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(world, res, what),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


EXEC_FLOP = 2 * (1024 * 512 + 512 * 256 + 256 * 512 + 768 * 256 + 512 * 128 + 128)      # chain kernel, per query


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mesh", action="store_true", help="skip the 512^3 mesh-latency leg")
    ap.add_argument("--no-encoders", action="store_true", help="skip the per-frame encoder leg")
    ap.add_argument("--res", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from pifu_b200 import config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    warmup = max(args.warmup, 3)
    res, what = workload(world)
    if args.res:
        res = args.res
    if res % world:
        raise SystemExit("the %d^3 lattice does not cut into %d slabs" % (res, world))

    torch.set_grad_enabled(False)
    # gate field (sigma 0.75, 2 % occupied): the parity gates of north_star are stated on this arithmetic
    netG, netMR, eng, calib, prob = build_nets(dev, saturated=False, with_prob=True)
    # host copies of the hot path's inputs (what filter_* leaves behind), pinned for the e2e leg
    feat_c_host = prob["feat_coarse"].pin_memory()
    feat_f_host = prob["feat_fine"].pin_memory()

    total = res ** 3
    per_rank = total // world
    id_b, id_e = rank * per_rank, (rank + 1) * per_rank
    slab = torch.empty(per_rank, device=dev, dtype=torch.float32)
    gathered = [torch.empty(per_rank, device=dev, dtype=torch.float32) for _ in range(world)] if (world > 1 and rank == 0) else None
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)

    def step():
        eng.eval_grid(2, res, calib[0], id_begin=id_b, id_end=id_e, out=slab)
        if world > 1:
            dist.gather(slab, gathered, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        step()
    barrier()
    l0 = eng.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    sampler.mark_begin()
    ev[0].record()
    for _ in range(args.steps):
        flush.fill_(0.0)                    # L2 flush between timed iterations (256 MiB write)
        step()
    ev[1].record()
    barrier()
    sampler.mark_end()
    ms = ev[0].elapsed_time(ev[1])
    launches = eng.launch_count() - l0 + args.steps     # + the flush fills
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = total / (ms_per_step * 1e-3)

    # ---- end to end through the reference-shaped API with HOST buffers: features H2D from pinned
    # memory + re-layout, the fused query, and the occupancy field back to pinned host memory
    field_host = torch.empty(per_rank, dtype=torch.float32).pin_memory()
    h2d = feat_c_host.numel() * 4 + feat_f_host.numel() * 4 + 16 * 4 + 16 * 8
    d2h = per_rank * 4

    def e2e_step():
        netG.im_feat_list = [feat_c_host.to(dev, non_blocking=True)]
        netMR.im_feat_list = [feat_f_host.to(dev, non_blocking=True)]
        eng.sync_features(0, netG.im_feat_list[-1])
        eng.sync_features(1, netMR.im_feat_list[-1])
        eng.eval_grid_host(2, res, calib[0], field_host, id_begin=id_b, id_end=id_e)

    e2e_step()
    barrier()
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev2[0].record()
    for _ in range(args.steps):
        e2e_step()
    ev2[1].record()
    barrier()
    t2 = torch.tensor([ev2[0].elapsed_time(ev2[1])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = total / (float(t2.item()) / args.steps * 1e-3)

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream
    roofline = cpu = parity = precision = None
    if rank == 0:
        peaks = load_peaks()
        eng.profile(True)
        eng.eval_grid(2, res, calib[0], id_begin=id_b, id_end=id_e, out=slab)
        n_c, chain_ms, chain_flops = eng.profile_read_kind(1)
        n_l, gemm_ms, gemm_flops = eng.profile_read_kind(0)
        eng.profile(False)
        alg_flop = config.FLOP_PER_QUERY_MR - 2 * (513 * 128 + 385)
        if n_c > 0:
            # dominant kernel = the lattice chain kernel (one launch per column chunk); the per-layer
            # kernel only computes the per-column constants.  `achieved` / `frac` count the FLOPs the tensor
            # cores EXECUTE (after the per-column constant folding); the reference's layer stack per query is
            # kept beside them as algorithmic_*.
            algorithmic = chain_flops / (chain_ms * 1e-3) / 1e12
            executed = algorithmic * EXEC_FLOP / alg_flop
            roofline = {"bound": "tensor", "kernel": "chain_kernel (tcgen05 CTA-pair MLP chain, activations on-chip)",
                        "achieved": executed, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": executed / peaks["tflops"],
                        "traffic": (ncu_traffic("chain_kernel") or {}).get("bytes_per_launch"),
                        "traffic_detail": ncu_traffic("chain_kernel"), "peak_source": peaks["source"], "launches_per_step": n_c,
                        "avg_launch_us": chain_ms * 1e3 / n_c, "share_of_step": chain_ms / ms_per_step,
                        "executed_flop_per_query": EXEC_FLOP, "algorithmic_flop_per_query": alg_flop,
                        "algorithmic_tflops": algorithmic, "algorithmic_frac": algorithmic / peaks["tflops"],
                        "column_constants": {"kernel": "gemm_tc_kernel", "launches_per_step": n_l, "ms_per_step": gemm_ms},
                        "note": "executed = what the tensor cores run after the per-column constant folding (SURVEY 7.3-4); "
                                "algorithmic = the reference's get_preds() layer stack per query (coarse L0-L2, fine L0-L3; "
                                "coarse L3/L4 only feed preds_low)"}
        else:
            achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
            roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 MLP layer)", "achieved": achieved,
                        "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                        "traffic": None, "peak_source": peaks["source"], "launches_per_step": n_l,
                        "avg_launch_us": gemm_ms * 1e3 / max(n_l, 1), "share_of_step": gemm_ms / ms_per_step,
                        "algorithmic_flop_per_query": alg_flop,
                        "note": "coarse L3/L4 (preds_low) are not on the get_preds() path and are skipped"}
        if not args.no_cpu_baseline:
            # the CPU port on a bounded sample of the SAME lattice; its values are the parity reference for the
            # field the timed step produced (this rank's slab)
            n_s = 128 ** 3
            qps, cores, npts, sec, ref_vals = cpu_port_queries_per_s(prob, calib, n_s, res=res, keep=True)
            cpu = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d-point strided sub-lattice of the %d^3 lattice, chunks of 100000, %.1f s" % (npts, res, sec)}
            ids = sample_ids(res, n_s)
            mine = (ids >= id_b) & (ids < id_e)
            # `slab` still holds this rank's field of the last evaluation (no collective here: only rank 0 is in this block)
            got = slab[torch.from_numpy(ids[mine] - id_b).to(dev)].cpu().numpy()
            parity = parity_block(got, ref_vals[mine], "timed step's field (fast arithmetic, gate field sigma 0.75) vs the CPU "
                                  "port on the strided sample%s" % ("" if world == 1 else ", rank 0's slab"))

    mesh = enc = stress = group = None
    del flush
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_mesh:
        del netMR, netG, eng
        # saturated field (x8 last-layer gain): what the octree prunes on and every mesh figure is taken on
        netG2, netMR2, eng2, calib2, prob2 = build_nets(dev, saturated=True, with_prob=True)
        # the three arithmetics on the dense 256^3 lattice of this field, each against the CPU port
        precision = {}
        ref_sat = None
        if not args.no_cpu_baseline:
            ref_sat = cpu_port_queries_per_s(prob2, calib2, 128 ** 3, res=res, keep=True)[4]
        ids_dev = torch.from_numpy(sample_ids(res, 128 ** 3)).to(dev)
        for mode, reps in (("fast", 5), ("hybrid", 5), ("split", 2)):
            eng2.set_precision(mode)
            eng2.eval_grid(2, res, calib2[0], out=slab)
            r0 = eng2.refined_points()
            evp = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            evp[0].record()
            for _ in range(reps):
                eng2.eval_grid(2, res, calib2[0], out=slab)
            evp[1].record()
            torch.cuda.synchronize(dev)
            d = {"queries_per_s": total / (evp[0].elapsed_time(evp[1]) / reps * 1e-3),
                 "ms_per_step": evp[0].elapsed_time(evp[1]) / reps}
            if mode == "hybrid":
                d["refined_fraction"] = (eng2.refined_points() - r0) / float(reps * total)
            if ref_sat is not None:
                d["parity"] = parity_block(slab[ids_dev].cpu().numpy(), ref_sat, "saturated field, %s arithmetic vs the CPU port" % mode)
            precision[mode] = d
        eng2.set_precision("fast")
        precision["note"] = ("dense %d^3 lattice of the saturated field (last layer x8); fast = one fp16 image per tensor-core operand; "
                             "split = fp16 + fp16 residual for features, activations and weights (3 products per layer, per-layer "
                             "kernels); hybrid = fast (chain kernel), then occupancies inside (0.02, 0.98) again in split" % res)
        group = group_norm_leg(dev, res)
        cpu_mesh = None
        if not args.no_cpu_baseline:
            cpu_mesh = mesh_cpu_baseline(prob2, calib2, 512)
        mesh = mesh_latency(netMR2, eng2, calib2, dev, 512, 3, cpu=cpu_mesh)
        if cpu_mesh is not None:
            mesh["cpu_baseline"] = {k: v for k, v in cpu_mesh.items() if k != "sdf"}
            mesh["cpu_baseline"]["kind"] = "port"
            mesh["cpu_baseline"]["what"] = ("create_grid + calib pre-transform, eval_grid_octree (numpy bookkeeping + torch-CPU query, "
                                            "num_samples 100000), marching cubes (oracle/mc_ref.c, single thread)")
            mesh["cpu_baseline_ms"] = cpu_mesh["total_ms"]
            mesh["octree"]["vs_cpu"]["mesh_cpu"] = [cpu_mesh["verts"], cpu_mesh["faces"]]
            mesh["octree"]["vs_cpu"]["mesh_gpu"] = [mesh["octree"]["verts"], mesh["octree"]["faces"]]
            for key in ("octree_hybrid", "octree_hybrid_narrow_band"):
                mesh[key]["vs_cpu"]["mesh_cpu"] = [cpu_mesh["verts"], cpu_mesh["faces"]]
                mesh[key]["vs_cpu"]["mesh_gpu"] = [mesh[key]["verts"], mesh[key]["faces"]]
        del cpu_mesh
        del netMR2, eng2, netG2
        torch.cuda.empty_cache()
        if not args.no_encoders:
            enc = encoder_leg(dev)

    if world > 1 and not args.no_mesh:
        del slab, gathered, netMR, netG, eng
        torch.cuda.empty_cache()
        _, netMR2, eng2, calib2 = build_mesh_problem(dev)        # seeded: identical on every rank
        mesh = mesh_latency_sharded(netMR2, calib2, dev, 512, 3)
        if world == 8:
            # BASELINE configs[4]: 1024^3 dense query + marching cubes on 8 GPUs (~1e9 points)
            stress = mesh_latency_sharded(netMR2, calib2, dev, 1024, 1, modes=("dense",))
        del netMR2, eng2
        torch.cuda.empty_cache()
        if not args.no_encoders:
            enc = encoder_leg(dev, frames=8, time_modes=False, world=world)

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": config_dict(world, res, what),
            "execution": {"parallelism": "slab%d" % world, "arithmetic": "f16 tensor-core operands, f32 accumulate and epilogues",
                          "path": "chain"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "precision": precision, "group_norm": group, "mesh_512": mesh, "stress_1024": stress, "encoders": enc,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
