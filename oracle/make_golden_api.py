"""Writes tests/golden/api_signatures.json: the parameter lists (names, order, defaults) of the reference's functions
and methods on the drop-in boundary, read with `inspect` from the UNMODIFIED reference imported from
/root/reference (authoring container only; same shims as make_golden.py).  tests/test_api_signatures_cpu.py holds the
package's mirror to this file, so "drops in under reconstruction.py" is checked mechanically and travels to the GPU box.
Usage:  python oracle/make_golden_api.py"""
import inspect
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, import_reference       # noqa: E402

BOUNDARY = {
    "mesh_util": ["create_grid", "batch_eval", "eval_grid", "eval_grid_octree", "reconstruction", "save_obj_mesh_with_color"],
    "PIFuNetwNML": ["__init__", "filter", "query", "get_preds", "get_im_feat", "calc_normal"],
    "PIFuMRNet": ["__init__", "filter_global", "filter_local", "query", "get_preds", "calc_normal"],
}


def describe(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        d = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append({"name": name, "kind": p.kind.name, "default": d})
    return out


def main():
    mesh_util, _, PIFuNetwNML, PIFuMRNet = import_reference()
    owners = {"mesh_util": mesh_util, "PIFuNetwNML": PIFuNetwNML, "PIFuMRNet": PIFuMRNet}
    sigs = {}
    for owner, names in BOUNDARY.items():
        for n in names:
            sigs["%s.%s" % (owner, n)] = describe(getattr(owners[owner], n))
    path = os.path.join(OUT, "api_signatures.json")
    with open(path, "w") as f:
        json.dump(sigs, f, indent=1, sort_keys=True)
    print("wrote", path, len(sigs), "signatures")


if __name__ == "__main__":
    main()
