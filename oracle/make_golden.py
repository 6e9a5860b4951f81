"""Generate tests/golden/*.npz by running the UNMODIFIED reference (authoring container only).

Usage:  python oracle/make_golden.py            (needs /root/reference; CPU, ~1-2 min)

The reference is imported from /root/reference with the shims recorded in SURVEY.md
Appendix A (a stub ``skimage`` module - scikit-image is not installed - and, on numpy
versions that lack it, the ``np.bool`` alias).  Its own ``PIFuNetwNML`` / ``PIFuMRNet`` / ``mesh_util`` produce every
number stored here; weights and feature maps come from ``pifu_b200.synthetic`` seeds so the
tests can regenerate bit-identical inputs without the reference.
"""
import copy
import hashlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PIFU_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.path.insert(0, REF)
    if not hasattr(np, "bool"):                       # mesh_util.py:134,136 (alias absent in numpy 1.24-1.26)
        np.bool = np.bool_
    sk = types.ModuleType("skimage")
    sk.measure = types.ModuleType("skimage.measure")
    sys.modules["skimage"] = sk
    sys.modules["skimage.measure"] = sk.measure      # mesh_util.py:10
    import mesh_util                                  # noqa
    from options import BaseOptions
    from PIFuNetwNML import PIFuNetwNML
    from PIFuMRNet import PIFuMRNet
    return mesh_util, BaseOptions, PIFuNetwNML, PIFuMRNet


def build_reference_nets(BaseOptions, PIFuNetwNML, PIFuMRNet, mlp_norm, mode="orthogonal"):
    """train.py:102-120 with the normal nets switched off (query path only)."""
    argv = sys.argv
    sys.argv = argv[:1]
    opt = BaseOptions().parse([])
    sys.argv = argv
    opt.mlp_norm = mlp_norm
    opt.use_front_normal = False
    opt.use_back_normal = False
    opt.hg_dim, opt.mlp_dim = opt.hg_dim_global, opt.mlp_dim_global
    opt.mlp_res_layers, opt.num_stack = opt.mlp_res_layers_global, opt.num_stack_global
    optG = copy.deepcopy(opt)
    netG = PIFuNetwNML(optG, mode)
    opt.num_stack, opt.hg_dim = opt.num_stack_local, opt.hg_dim_local
    opt.mlp_dim, opt.mlp_res_layers = opt.mlp_dim_local, opt.mlp_res_layers_local
    netMR = PIFuMRNet(opt, netG, mode)
    netG.eval()                                      # reconstruction.py:288-289
    return netG, netMR


def load_problem(netG, netMR, prob):
    netG.mlp.load_state_dict(prob["coarse"], strict=False)
    netMR.mlp.load_state_dict(prob["fine"], strict=False)
    netG.im_feat_list = [prob["feat_coarse"]]        # PIFuNetwNML.py:94-97
    netMR.im_feat_list = [prob["feat_fine"]]         # PIFuMRNet.py:114-117


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def soft_ellipsoid(points):
    """Analytic eval_func (float32 result like mesh_util.py:74); +,-,*,/,sqrt only."""
    x, y, z = points[0], points[1], points[2]
    r = np.sqrt((x / 0.35) ** 2 + (y / 0.8) ** 2 + (z / 0.3) ** 2)
    return np.clip(0.5 + 2.0 * (1.0 - r), 0.0, 1.0).astype(np.float32)


def ripple_field(points):
    """Rougher analytic field: products of the coordinates, many narrow skip decisions."""
    x, y, z = points[0], points[1], points[2]
    v = 0.5 + 0.6 * (x * y - z * z) + 0.25 * (x * x * x - y * z)
    return np.clip(v, 0.0, 1.0).astype(np.float32)


def main():
    from pifu_b200 import synthetic as syn
    mesh_util, BaseOptions, PIFuNetwNML, PIFuMRNet = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_grad_enabled(False)

    # ---------------------------------------------------------------- query goldens
    netG, netMR = build_reference_nets(BaseOptions, PIFuNetwNML, PIFuMRNet, "none")
    prob = syn.make_problem(bias_std=0.01)
    load_problem(netG, netMR, prob)
    calib = syn.default_calib()
    pilot = syn.random_points(20000, syn.SEED_PILOT, -1.0, 1.0)
    netMR.query(pilot, calib)
    pilot_preds = netMR.get_preds().numpy().copy()
    syn.calibrate_last_layer(prob["fine"], 3, pilot_preds)
    load_problem(netG, netMR, prob)

    pts = syn.random_points(2048)
    out = {"pilot_preds": pilot_preds[0, 0, :4096].astype(np.float32)}

    netG.query(pts, calib)
    out["coarse_preds"] = netG.get_preds().numpy().copy()
    out["coarse_phi_128"] = netG.phi[:, :, :128].numpy().copy()

    netMR.query(pts, calib)
    out["mr_preds"] = netMR.get_preds().numpy().copy()
    out["mr_preds_low"] = netMR.preds_low.numpy().copy()
    out["mr_preds_interm"] = netMR.preds_interm.numpy().copy()

    calib2 = syn.scaled_calib()
    netMR.query(pts, calib2)
    out["mr_preds_scaled_calib"] = netMR.get_preds().numpy().copy()

    # multi-crop call form (PIFuMRNet.py:119-186 with calib_global given, B2 = 1)
    netMR.query(pts[:, None], calib2[:, None], calib)
    out["mr_preds_local_global"] = netMR.get_preds().numpy().copy()

    # calc_normal (PIFuMRNet.py:188-243)
    netMR.calc_normal(pts[:, None, :, :256], calib[:, None], calib, delta=0.001)
    out["mr_nmls_256"] = netMR.nmls.numpy().copy()

    # dense lattice through the reference driver (mesh_util.py:59-80, eval_grid)
    res = 16
    coords, mat = mesh_util.create_grid(res, res, res)
    cinv = np.linalg.inv(calib[0].numpy())
    c = coords.reshape(3, -1).T
    c = np.matmul(np.concatenate([c, np.ones((c.shape[0], 1))], 1), cinv.T)[:, :3]
    coords = c.T.reshape(3, res, res, res)

    def eval_func(points):
        samples = torch.from_numpy(np.expand_dims(points, 0)).float()
        netMR.query(samples, calib)
        return netMR.get_preds()[0][0].numpy()
    out["mr_grid16"] = mesh_util.eval_grid(coords, eval_func, num_samples=1000)
    out["grid16_coords_sha"] = np.frombuffer(bytes.fromhex(sha(coords)), dtype=np.uint8)
    out["grid16_mat"] = mat

    # octree through the reference with the real net on the saturated field, 64^3 / init 16
    syn.saturate(prob["fine"], 3)
    load_problem(netG, netMR, prob)
    res = 64
    coords, mat = mesh_util.create_grid(res, res, res)
    c = coords.reshape(3, -1).T
    c = np.matmul(np.concatenate([c, np.ones((c.shape[0], 1))], 1), cinv.T)[:, :3]
    coords = c.T.reshape(3, res, res, res)
    f = mesh_util.eval_grid_octree(coords, eval_func, init_resolution=16, num_samples=100000)
    out["mr_octree64_init16"] = f.astype(np.float32)
    out["mr_octree64_init16_sha_f64"] = np.frombuffer(bytes.fromhex(sha(f)), dtype=np.uint8)

    # perspective projection (PIFuMRNet's default-mode typo selects it, BasePIFuNet.py:79)
    syn.saturate(prob["fine"], 3, 1.0 / syn.SATURATE_GAIN)
    load_problem(netG, netMR, prob)
    netGp, _ = build_reference_nets(BaseOptions, PIFuNetwNML, PIFuMRNet, "none", "perspective")
    netGp.mlp.load_state_dict(prob["coarse"], strict=False)
    netGp.im_feat_list = [prob["feat_coarse"]]
    ppts = pts.clone()
    ppts[:, 2, :] = ppts[:, 2, :] * 0.25 + 2.0
    netGp.query(ppts, calib)
    out["coarse_preds_perspective"] = netGp.get_preds().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "query_none.npz"), **out)

    # ---------------------------------------------------------------- group-norm golden
    netGg, netMRg = build_reference_nets(BaseOptions, PIFuNetwNML, PIFuMRNet, "group")
    load_problem(netGg, netMRg, prob)
    netMRg.query(pts, calib)
    np.savez_compressed(os.path.join(OUT, "query_group.npz"),
                        mr_preds=netMRg.get_preds().numpy().copy(),
                        mr_preds_low=netMRg.preds_low.numpy().copy())

    # ---------------------------------------------------------------- analytic octree
    oct_out = {}
    for name, fn in (("ellipsoid", soft_ellipsoid), ("ripple", ripple_field)):
        for res, init in ((64, 8), (64, 16), (96, 12), (128, 32), (128, 64)):
            coords, _ = mesh_util.create_grid(res, res, res)
            calls = []

            def counting(points, fn=fn, calls=calls):
                calls.append(points.shape[1])
                return fn(points)
            f = mesh_util.eval_grid_octree(coords, counting, init_resolution=init,
                                           num_samples=50000)
            key = "%s_%d_%d" % (name, res, init)
            oct_out[key + "_sha"] = np.frombuffer(bytes.fromhex(sha(f)), dtype=np.uint8)
            oct_out[key + "_evaluated"] = np.array(sum(calls), dtype=np.int64)
            oct_out[key + "_sum"] = np.array(f.sum(), dtype=np.float64)
            oct_out[key + "_probe"] = f[::7, ::5, ::3].astype(np.float64)
    np.savez_compressed(os.path.join(OUT, "octree_analytic.npz"), **oct_out)

    for fn_ in sorted(os.listdir(OUT)):
        print(fn_, os.path.getsize(os.path.join(OUT, fn_)))


if __name__ == "__main__":
    main()
