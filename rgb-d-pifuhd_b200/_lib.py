"""ctypes binding of libpifu_b200.so (C ABI in include/pifu_b200.h).

There is no fallback: if the library is missing it is built with nvcc when a toolchain is
present, otherwise importing a compute entry point raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpifu_b200.so")
_lib = None

c_float_p = ctypes.POINTER(ctypes.c_float)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_ll_p = ctypes.POINTER(ctypes.c_longlong)
VP = ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol include/pifu_b200.h declares
SIGNATURES = {
    "pifu_last_error": (ctypes.c_char_p, []),
    "pifu_abi_version": (ctypes.c_int, []),
    "pifu_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(VP)]),
    "pifu_destroy": (None, [VP]),
    "pifu_set_mlp": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, c_int_p, ctypes.c_int, c_int_p,
                                    ctypes.c_int, ctypes.POINTER(VP), ctypes.POINTER(VP), VP]),
    "pifu_set_features": (ctypes.c_int, [VP, ctypes.c_int, VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, VP]),
    "pifu_set_options": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_float, ctypes.c_float]),
    "pifu_set_gemm_impl": (ctypes.c_int, [VP, ctypes.c_int]),
    "pifu_set_precision": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]),
    "pifu_refined_points": (ctypes.c_longlong, [VP]),
    "pifu_set_chunk_tiles": (ctypes.c_int, [VP, ctypes.c_int]),
    "pifu_set_mlp_norm": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, ctypes.c_double, VP, VP, VP]),
    "pifu_query": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, VP, ctypes.c_longlong, ctypes.c_longlong,
                                  c_float_p, c_float_p, VP, VP, VP, VP]),
    "pifu_eval_grid": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_longlong, ctypes.c_longlong, c_float_p, c_double_p, VP, VP]),
    "pifu_eval_lattice_ids": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             VP, ctypes.c_longlong, c_float_p, c_double_p, VP, VP]),
    "pifu_eval_grid_octree": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_double, c_float_p, c_double_p, VP, VP, c_ll_p, ctypes.c_int, VP]),
    "pifu_octree_begin": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, VP]),
    "pifu_octree_frontier": (ctypes.c_int, [VP, c_ll_p, ctypes.POINTER(VP), c_int_p, VP]),
    "pifu_octree_commit": (ctypes.c_int, [VP, VP, VP]),
    "pifu_octree_commit64": (ctypes.c_int, [VP, VP, VP]),
    "pifu_octree_export": (ctypes.c_int, [VP, VP, VP, VP]),
    "pifu_octree_begin_slab": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, VP]),
    "pifu_octree_commit_pairs": (ctypes.c_int, [VP, VP, VP, ctypes.c_longlong, VP]),
    "pifu_octree_set_frontier_planes": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int]),
    "pifu_octree_field32": (ctypes.c_int, [VP, ctypes.POINTER(VP), c_int_p, c_int_p]),
    "pifu_mc_count": (ctypes.c_int, [VP, VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                     c_ll_p, c_ll_p, VP]),
    "pifu_mc_count_slab": (ctypes.c_int, [VP, VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          c_ll_p, c_ll_p, c_ll_p, VP]),
    "pifu_mc_emit": (ctypes.c_int, [VP, VP, VP, VP, VP, VP]),
    "pifu_mc_extract": (ctypes.c_int, [VP, VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, VP, VP, VP, VP, ctypes.c_longlong,
                                       ctypes.c_longlong, VP, VP]),
    "pifu_write_obj": (ctypes.c_int, [ctypes.c_char_p, VP, VP, ctypes.c_longlong, VP, ctypes.c_longlong]),
    "pifu_obj_counts": (ctypes.c_int, [ctypes.c_char_p, VP]),
    "pifu_read_obj": (ctypes.c_int, [ctypes.c_char_p, VP, VP, VP, ctypes.c_longlong, ctypes.c_longlong]),
    "pifu_bn_relu_f32": (ctypes.c_int, [VP, VP, VP, VP, VP, ctypes.c_double, ctypes.c_int, VP, ctypes.c_longlong,
                                        ctypes.c_int, ctypes.c_longlong, VP]),
    "pifu_cat3_add_f32": (ctypes.c_int, [VP, VP, VP, VP, VP, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                                         ctypes.c_longlong, VP]),
    "pifu_sample_image": (ctypes.c_int, [VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, VP, ctypes.c_longlong, ctypes.c_longlong,
                                         c_float_p, ctypes.c_int, VP, VP]),
    "pifu_mesh_clean": (ctypes.c_int, [VP, VP, VP, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, VP, VP, VP, c_ll_p, VP]),
    "pifu_launch_count": (ctypes.c_longlong, [VP]),
    "pifu_profile_enable": (ctypes.c_int, [VP, ctypes.c_int]),
    "pifu_profile_read": (ctypes.c_int, [VP, c_ll_p, c_double_p, c_double_p]),
    "pifu_profile_read_kind": (ctypes.c_int, [VP, ctypes.c_int, c_ll_p, c_double_p, c_double_p]),
    "pifu_set_chain": (ctypes.c_int, [VP, ctypes.c_int]),
    "pifu_chain_ready": (ctypes.c_int, [VP]),
    "pifu_debug_gemm": (ctypes.c_int, [VP, VP, VP, VP, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, VP, VP]),
}


ABI_VERSION = 3          # pifu_abi_version() of the library these signatures describe


class PifuError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def load():
    """Load (building first if needed) and type the library.  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    stale = not os.path.exists(_LIB_PATH) or not _build.up_to_date()
    if stale:
        # missing, or built from other sources than the ones in the tree (csrc / header digest in the stamp)
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(_LIB_PATH):
                raise PifuError("libpifu_b200.so is missing and could not be built (%s). "
                                "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root; "
                                "there is no CPU fallback." % (e,))
            raise PifuError("libpifu_b200.so does not match the sources in the tree and could not be rebuilt (%s)" % (e,))
    lib = ctypes.CDLL(_LIB_PATH)
    lib.pifu_abi_version.restype = ctypes.c_int
    if lib.pifu_abi_version() != ABI_VERSION:
        raise PifuError("libpifu_b200.so has ABI version %d, this binding expects %d: rebuild it"
                        % (lib.pifu_abi_version(), ABI_VERSION))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PifuError(load().pifu_last_error().decode("utf-8", "replace"))
