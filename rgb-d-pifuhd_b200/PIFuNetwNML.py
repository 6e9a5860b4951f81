"""Coarse pixel-aligned implicit function, drop-in for the reference's `PIFuNetwNML.py`.

`filter` stays PyTorch (it runs once per image and only orchestrates the caller's encoder);
`query` / `get_preds` / `calc_normal` run in libpifu_b200.so."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .BasePIFuNet import BasePIFuNet, _not_hot_path
from .MLP import MLP
from .engine import get_engine


class PIFuNetwNML(BasePIFuNet):
    """Same constructor as the reference (`PIFuNetwNML.py:19-71`) plus ``image_filter``: the
    hourglass encoder (reference `Filter`, or any module returning ``(feature_list, normx)``)
    is supplied by the caller - it is not part of the replaced path."""

    def __init__(self, opt, projection_mode="orthogonal", criteria=None, image_filter=None):
        super().__init__(projection_mode=projection_mode, criteria=criteria)
        self.name = "hg_pifu"
        self.opt = opt
        self.image_filter = image_filter
        self.mlp = MLP(filter_channels=opt.mlp_dim, merge_layer=opt.merge_layer,
                       res_layers=opt.mlp_res_layers, norm=opt.mlp_norm, last_op=nn.Sigmoid())
        for f in self.mlp.filters:                                  # net_util.py:13-25
            nn.init.normal_(f.weight, 0.0, 0.02)
            nn.init.constant_(f.bias, 0.0)
        self.im_feat_list = []
        self.tmpx = None
        self.normx = None
        self.phi = None
        self.intermediate_preds_list = []
        self.netF = None
        self.netB = None
        self.nmlF = None
        self.nmlB = None

    # ------------------------------------------------------------------ encoder (PyTorch, once per image)
    def filter(self, images):
        """`PIFuNetwNML.py:73-97`: optional front/back normal nets, concat, encoder."""
        if self.image_filter is None:
            raise RuntimeError("no image_filter was given to PIFuNetwNML; assign im_feat_list directly "
                               "or pass the reference's Filter module")
        extra = []
        with torch.no_grad():
            if self.netF is not None:
                self.nmlF = self.netF.forward(images).detach()
                extra.append(self.nmlF)
            if self.netB is not None:
                self.nmlB = self.netB.forward(images).detach()
                extra.append(self.nmlB)
        if extra:
            nmls = torch.cat(extra, 1)
            if nmls.shape[2:] != images.shape[2:]:
                nmls = F.interpolate(nmls, size=images.shape[2:], mode="bilinear", align_corners=True)
            images = torch.cat([images, nmls], 1)
        self.im_feat_list, self.normx = self.image_filter(images)
        if not self.training:
            self.im_feat_list = [self.im_feat_list[-1]]

    # ------------------------------------------------------------------ fused query
    def _engine_for(self, points):
        eng = get_engine(points.device)
        eng.set_options(self.is_perspective, self.opt.loadSize, self.opt.z_size)
        eng.sync_mlp(0, self.mlp, id(self))
        return eng

    def query(self, points, calibs, transforms=None, labels=None, update_pred=True, update_phi=True):
        """`PIFuNetwNML.py:99-141`.  points [B, 3, N], calibs [B, 4, 4] (or [B, 3, 4])."""
        if transforms is not None:
            _not_hot_path("screen-space `transforms`")
        if labels is not None:
            _not_hot_path("training supervision (`labels`)")
        if len(self.im_feat_list) != 1:
            _not_hot_path("train-mode query over %d intermediate feature maps" % len(self.im_feat_list))
        eng = self._engine_for(points)
        feat = self.im_feat_list[-1]
        cphi = self.mlp.filter_channels[self.mlp.merge_layer + 1]
        preds, phis = [], []
        for b in range(points.shape[0]):
            eng.sync_features(0, feat[b:b + 1])
            p, _, phi = eng.query(1, points[b], calibs[b], calibs[b], want_phi=cphi if update_phi else 0)
            preds.append(p[None, None])
            phis.append(phi[None] if phi is not None else None)
        pred = torch.cat(preds, 0)
        if update_phi:
            self.phi = torch.cat(phis, 0)
        if update_pred:
            self.intermediate_preds_list = [pred]
            self.preds = pred

    def calc_normal(self, points, calibs, transforms=None, labels=None, delta=0.1):
        """`PIFuNetwNML.py:181-220`: forward differences of the (un-masked) coarse occupancy."""
        if transforms is not None or labels is not None:
            _not_hot_path("`transforms` / `labels`")
        shifted = [points] + [points.clone() for _ in range(3)]
        for a in range(3):
            shifted[a + 1][:, a, :] += delta
        pall = torch.stack(shifted, 3).reshape(points.shape[0], 3, -1)
        # the reference applies no in-bounds mask here (`PIFuNetwNML.py:211`)
        eng = self._engine_for(points)
        preds = []
        for b in range(points.shape[0]):
            eng.sync_features(0, self.im_feat_list[-1][b:b + 1])
            preds.append(eng.query(1, pall[b], calibs[b], calibs[b], no_mask=True)[0][None, None])
        pred = torch.cat(preds, 0).view(points.shape[0], 1, -1, 4)
        d = [pred[:, :, :, a + 1] - pred[:, :, :, 0] for a in range(3)]
        self.nml = F.normalize(-torch.cat(d, 1), dim=1, eps=1e-8)

    def get_im_feat(self):
        return self.im_feat_list[-1]

    def loadFromPIFu(self, net):
        _not_hot_path("checkpoint surgery (loadFromPIFu)")

    def forward(self, *a, **k):
        _not_hot_path("the training forward")
