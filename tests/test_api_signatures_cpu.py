"""The drop-in boundary, checked mechanically: every function / method of the reference that
`reconstruction.py` / `eval.py` call on this path keeps its parameter names, their order and their defaults in the
package's mirror.  The reference's side is tests/golden/api_signatures.json, read with `inspect` from the unmodified
reference by oracle/make_golden_api.py.  The mirror may ADD parameters only after the reference's and only with defaults
(keyword-only ones included), so every call the reference's callers make binds the same way."""
import inspect
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN


def _reference():
    with open(os.path.join(GOLDEN, "api_signatures.json")) as f:
        return json.load(f)


def _owners():
    from pifu_b200 import mesh_util
    from pifu_b200.PIFuMRNet import PIFuMRNet
    from pifu_b200.PIFuNetwNML import PIFuNetwNML
    return {"mesh_util": mesh_util, "PIFuNetwNML": PIFuNetwNML, "PIFuMRNet": PIFuMRNet}


def _same_default(ours, ref_repr):
    if ref_repr is None:
        return ours is inspect.Parameter.empty
    if ours is inspect.Parameter.empty:
        return False
    if isinstance(ours, np.ndarray):                                    # create_grid's b_min / b_max
        return repr(ours) == ref_repr
    if ref_repr.startswith("{'occ': MSELoss()"):                         # the criteria dict: same keys, same loss class
        return isinstance(ours, dict) and list(ours) == ["occ"] and type(ours["occ"]).__name__ == "MSELoss"
    return repr(ours) == ref_repr


@pytest.mark.parametrize("name", sorted(_reference()))
def test_signature_matches_reference(name):
    owner, attr = name.split(".")
    fn = getattr(_owners()[owner], attr)
    ours = list(inspect.signature(fn).parameters.values())
    ref = _reference()[name]
    assert len(ours) >= len(ref), "%s lost parameters: %s" % (name, [p.name for p in ours])
    for p, r in zip(ours, ref):
        assert p.name == r["name"], "%s: parameter %r where the reference has %r" % (name, p.name, r["name"])
        assert p.kind.name == r["kind"], "%s.%s: kind %s, reference %s" % (name, p.name, p.kind.name, r["kind"])
        assert _same_default(p.default, r["default"]), "%s.%s: default %r, reference %s" % (name, p.name, p.default, r["default"])
    for p in ours[len(ref):]:                                            # additions must not change how old calls bind
        assert p.default is not inspect.Parameter.empty or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL), \
            "%s: added parameter %r has no default" % (name, p.name)


def test_attributes_the_callers_read():
    """What `reconstruction.py` reads off the nets (`:37-40` netF / netB / nmlF / nmlB, `:69` nmls, `:113,165,182`
    projection, `:285-292` nn.Module life cycle) exists on the mirror with the reference's meaning."""
    import torch
    from torch import nn
    from pifu_b200 import config
    from pifu_b200.BasePIFuNet import orthogonal, perspective
    from pifu_b200.PIFuMRNet import PIFuMRNet
    from pifu_b200.PIFuNetwNML import PIFuNetwNML
    netG = PIFuNetwNML(config.coarse_opt(), "orthogonal", image_filter=None)
    netMR = PIFuMRNet(config.fine_opt(), netG, "orthogonal", image_filter=None)
    assert isinstance(netG, nn.Module) and isinstance(netMR, nn.Module) and netMR.netG is netG
    for a in ("netF", "netB", "nmlF", "nmlB", "phi", "im_feat_list", "opt", "projection", "preds", "mlp", "training"):
        assert hasattr(netG, a), a
    for a in ("nmls", "im_feat_list", "opt", "projection", "preds", "preds_interm", "preds_low", "mlp", "training"):
        assert hasattr(netMR, a), a
    assert netMR.projection is orthogonal and netG.projection is orthogonal
    assert PIFuMRNet(config.fine_opt(), netG, image_filter=None).projection is perspective   # the reference's default typo (`PIFuMRNet.py:22`)
    assert list(netG.criteria) == ["occ"]
    netMR.eval()
    assert not netMR.training and not netG.training
    sd = netMR.state_dict()                                              # load_state_dict round trip (`reconstruction.py:290-292`)
    netMR.load_state_dict(sd)
    assert any(k.startswith("netG.mlp.filters.") for k in sd) and any(k.startswith("mlp.filters.") for k in sd)
    assert torch.is_tensor(sd["mlp.filters.0.weight"])
