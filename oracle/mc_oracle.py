"""CPU ORACLE (test infrastructure): ctypes face of oracle/mc_ref.c (sequential marching cubes that derives
every cell's triangles at run time from the rule in rgb-d-pifuhd_b200/tools/gen_mc_tables.py - no lookup table
shared with the product).  Built with gcc into oracle/_build/ by ``build()``; see mc_ref.c for what it follows."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmc_ref.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "mc_ref.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    # -ffp-contract=off: no fused multiply-add, so the float64 arithmetic is the plain IEEE
    # sequence the CUDA kernel spells out with __dmul_rn/__dadd_rn
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.mc_ref_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
        _lib.mc_ref_run.restype = ctypes.c_int
        _lib.mc_ref_run_slab.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib.mc_ref_run_slab.restype = ctypes.c_int
        _lib.mc_ref_ghost_verts.restype = ctypes.c_longlong
        _lib.mc_ref_cell.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
        _lib.mc_ref_cell.restype = ctypes.c_int
    return _lib


def cell_triangles(corner_values, level):
    """The rule applied to ONE cell: eight corner values (corner order of the rule) -> flat list of edge ids."""
    v = np.ascontiguousarray(corner_values, dtype=np.float64)
    out = np.zeros(30, dtype=np.int32)
    n = _load().mc_ref_cell(v.ctypes.data, float(level), out.ctypes.data)
    return [int(x) for x in out[:3 * n]]


def marching_cubes(volume, level):
    """volume [n0,n1,n2] (cast to float32 like skimage does) -> verts f64 [V,3] (volume-index
    coordinates), faces i32 [F,3], normals f32 [V,3], values f32 [V], cases u8 per cell.
    Raises ValueError when the level is outside the data range / no surface (skimage behaviour
    that `mesh_util.py:94-96` turns into -1)."""
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    if vol.ndim != 3 or min(vol.shape) < 2:
        raise ValueError("Input volume should be a 3D numpy array with at least 2 points per axis")
    if level < vol.min() or level > vol.max():
        raise ValueError("Surface level must be within volume data range.")
    lib = _load()
    if lib.mc_ref_run(vol.ctypes.data, vol.shape[0], vol.shape[1], vol.shape[2], float(level)) != 0:
        raise MemoryError("mc_ref_run")
    nv, nf, nc = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong()
    lib.mc_ref_sizes(ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(nc))
    verts = np.empty((nv.value, 3), np.float64)
    faces = np.empty((nf.value, 3), np.int32)
    normals = np.empty((nv.value, 3), np.float32)
    values = np.empty((nv.value,), np.float32)
    cases = np.empty((nc.value,), np.uint8)
    lib.mc_ref_copy(verts.ctypes.data_as(ctypes.c_void_p), faces.ctypes.data_as(ctypes.c_void_p),
                    normals.ctypes.data_as(ctypes.c_void_p), values.ctypes.data_as(ctypes.c_void_p),
                    cases.ctypes.data_as(ctypes.c_void_p))
    if nv.value == 0:
        raise RuntimeError("No surface found at the given iso value.")
    return verts, faces, normals, values, cases.reshape(vol.shape[0] - 1, vol.shape[1] - 1, vol.shape[2] - 1)


def marching_cubes_slab(volume, level, i_global0, global_n0, cell_layers, ghost):
    """Slab form with the signature of Engine.marching_cubes_slab (numpy in / numpy out):
    -> (verts, faces, normals, values, ghost_verts).  An empty slab gives zero-length arrays."""
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    lib = _load()
    if lib.mc_ref_run_slab(vol.ctypes.data, vol.shape[0], vol.shape[1], vol.shape[2], float(level),
                           int(i_global0), int(global_n0), int(cell_layers), 1 if ghost else 0) != 0:
        raise MemoryError("mc_ref_run_slab")
    nv, nf, nc = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong()
    lib.mc_ref_sizes(ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(nc))
    verts = np.empty((nv.value, 3), np.float64)
    faces = np.empty((nf.value, 3), np.int32)
    normals = np.empty((nv.value, 3), np.float32)
    values = np.empty((nv.value,), np.float32)
    lib.mc_ref_copy(verts.ctypes.data_as(ctypes.c_void_p), faces.ctypes.data_as(ctypes.c_void_p),
                    normals.ctypes.data_as(ctypes.c_void_p), values.ctypes.data_as(ctypes.c_void_p), None)
    return verts, faces, normals, values, int(lib.mc_ref_ghost_verts())
