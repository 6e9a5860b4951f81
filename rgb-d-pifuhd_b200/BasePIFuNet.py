"""Base class mirroring the attributes callers of the reference read (`BasePIFuNet.py:67-148`)."""
import torch
import torch.nn as nn


def _not_hot_path(what):
    raise NotImplementedError("%s is outside the reconstruction hot path this package replaces" % what)


def index(feat, uv):
    """`BasePIFuNet.py:11-23`: bilinear samples [B, C, N] of `feat` [B, C, H, W] at `uv` [B, 2, N] in [-1, 1]
    (align_corners=True, zeros outside).  On the query path this is fused into gather.cu; the function stays
    for callers that sample an *image*, e.g. the vertex colours of `gen_mesh_imgColor`
    (`reconstruction.py:110-116`)."""
    grid = uv.transpose(1, 2).unsqueeze(2)
    return torch.nn.functional.grid_sample(feat, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[:, :, :, 0]


def orthogonal(points, calib, transform=None):
    """`BasePIFuNet.py:25-43` - kept for callers that project by hand (`reconstruction.py:113`)."""
    if transform is not None:
        _not_hot_path("screen-space `transform`")
    return torch.baddbmm(calib[:, :3, 3:4], calib[:, :3, :3], points)


def perspective(points, calib, transform=None):
    """`BasePIFuNet.py:45-65`."""
    if transform is not None:
        _not_hot_path("screen-space `transform`")
    homo = torch.baddbmm(calib[:, :3, 3:4], calib[:, :3, :3], points)
    return torch.cat([homo[:, :2, :] / homo[:, 2:3, :], homo[:, 2:3, :]], 1)


def check_batch_statistics(mlp, n_items):
    """A BatchNorm1d MLP left in train mode takes its statistics over ALL items of the call ([B, C, N] -> per channel over
    B x N, `MLP.py:36-41`); the fused path evaluates the items of a batch one after the other, each with its own
    statistics - the reference's numbers for a batch of one (what `reconstruction()` passes), not for more.  GroupNorm
    normalises per item, so it is unaffected."""
    if n_items > 1 and getattr(mlp, "norm", "none") == "batch" and mlp.training:
        _not_hot_path("a train-mode BatchNorm1d MLP queried with a batch of %d items (statistics across the batch)" % n_items)


class BasePIFuNet(nn.Module):
    def __init__(self, projection_mode="orthogonal", criteria=None):
        super().__init__()
        self.name = "base"
        self.criteria = criteria
        # any string other than 'orthogonal' selects perspective (`BasePIFuNet.py:79`)
        self.projection_mode = projection_mode
        self.projection = orthogonal if projection_mode == "orthogonal" else perspective
        self.preds = None
        self.labels = None
        self.nmls = None
        self.labels_nml = None
        self.preds_surface = None

    @property
    def is_perspective(self):
        return self.projection is not orthogonal

    def get_preds(self):
        """`BasePIFuNet.py:136-142`."""
        return self.preds

    def get_error(self, gamma=None):
        _not_hot_path("training loss")
