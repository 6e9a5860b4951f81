"""Hot-path option fields.

The reference reads its model dimensions from an ``argparse.Namespace`` built by
``options.py:6-216``; ``train.py:102-120`` copies the ``*_global`` values into the
coarse net's namespace and the ``*_local`` values into the fine net's.  Only the
fields the reconstruction path consumes are reproduced here, with the reference's
names, so the nets in this package accept either the reference's own namespace or
one made by :func:`coarse_opt` / :func:`fine_opt`.

One default differs on purpose: ``mlp_norm`` is ``'none'`` here, ``'group'`` in
``options.py:95``.  GroupNorm's statistics run over all points of a ``query()`` call
(SURVEY 7.3-1), so a normalised MLP can neither be fused on-chip nor sharded; ``'none'``
is a first-class reference configuration (``MLP.py:66-67``) and the one the headline
numbers are quoted on.  A namespace parsed by the reference's own ``options.py``
carries ``'group'`` and is honoured: those nets take the per-layer kernels with
call-wide statistics (norm.cu); bench.py reports that configuration's rate beside
the headline (``group_norm``).
"""
from argparse import Namespace

# options.py:96,102,108 / :98,104 (defaults), :18 loadSize, :73 z_size, :152 loadSizeBig
_COMMON = dict(loadSize=1024, loadSizeBig=1024, z_size=200.0, merge_layer=2,
               mlp_norm="none", norm="batch", hg_depth=2, hg_down="ave_pool",
               train_full_pifu=False, use_front_normal=False, use_back_normal=False,
               no_intermediate_loss=False)


def coarse_opt(**over):
    """Namespace of the coarse net (`train.py:102-108`)."""
    d = dict(_COMMON, mlp_dim=[257, 1024, 512, 256, 128, 1], mlp_res_layers=[2, 3, 4],
             num_stack=4, hg_dim=256)
    d.update(over)
    return Namespace(**d)


def fine_opt(**over):
    """Namespace of the fine net (`train.py:115-120`)."""
    d = dict(_COMMON, mlp_dim=[272, 512, 256, 128, 1], mlp_res_layers=[1, 2],
             num_stack=1, hg_dim=16)
    d.update(over)
    return Namespace(**d)


# FLOP per query used for the tensor roofline:
# 2 * MACs of the conv1d(k=1) stacks.
COARSE_MACS = 257 * 1024 + 1024 * 512 + 769 * 256 + 513 * 128 + 385
FINE_MACS = 272 * 512 + 784 * 256 + 528 * 128 + 128
FLOP_PER_QUERY_MR = 2 * (COARSE_MACS + FINE_MACS)          # 2 916 098
FLOP_PER_QUERY_COARSE = 2 * COARSE_MACS                    # 2 100 738
