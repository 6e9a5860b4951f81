"""Coarse pixel-aligned implicit function, drop-in for the reference's `PIFuNetwNML.py`.

`filter` stays PyTorch (it runs once per image: hourglass `Filter` + optional normal nets, executed
by encoders.EncoderRunner); `query` / `get_preds` / `calc_normal` run in libpifu_b200.so."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import encoders
from .BasePIFuNet import BasePIFuNet, _not_hot_path, check_batch_statistics
from .MLP import MLP
from .engine import get_engine


class _AsEncoder(nn.Module):
    """A tensor -> tensor module (the normal nets) behind the encoder calling convention."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x):
        return [self.net(x)], None


class EncoderHost:
    """Shared by the coarse and the fine net: how their PyTorch encoders are executed.
    `encoder_opts` (channels_last / precision / graph, see encoders.EncoderRunner) may be changed
    before the first `filter*` call or followed by `reset_encoder_runners()`.  The default (TF32
    convolutions, NCHW, graph replay) is the arithmetic stock PyTorch gives the reference on this GPU
    (`torch.backends.cudnn.allow_tf32` defaults to True); `precision='fp32'` is IEEE fp32 (3x slower),
    `'bf16'` autocast trades ~1e-2 relative feature error for less element-wise traffic."""

    encoder_opts = None

    def _runner(self, key, module):
        if self.encoder_opts is None:
            self.encoder_opts = dict(channels_last=False, precision="tf32", graph=True)
        runners = self.__dict__.setdefault("_enc_runners", {})
        r = runners.get(key)
        if r is None or r.module is not module:
            r = encoders.EncoderRunner(module, **self.encoder_opts)
            runners[key] = r
        return r

    def reset_encoder_runners(self):
        self.__dict__.pop("_enc_runners", None)

    def _run_normal_net(self, key, net, images):
        wraps = self.__dict__.setdefault("_enc_wrap", {})
        w = wraps.get(key)
        if w is None or w.net is not net:
            w = _AsEncoder(net)
            wraps[key] = w
        w.train(net.training)
        return self._run_encoder(key, w, images, True)[0][0].detach()

    def _run_encoder(self, key, module, images, last_only):
        """(feature_list, normx) of `module(images)`; tuned execution only for inference on a CUDA device."""
        if images.is_cuda and not torch.is_grad_enabled() and not module.training:
            return self._runner(key, module)(images, last_only=last_only)
        feats, normx = module(images)
        return ([feats[-1]] if last_only else feats), normx


class PIFuNetwNML(BasePIFuNet, EncoderHost):
    """Same constructor as the reference (`PIFuNetwNML.py:19-71`) plus ``image_filter``: 'auto' builds
    the hourglass `Filter` the reference builds (`:40-41`; None when `opt` has no encoder fields); a
    module returning ``(feature_list, normx)`` - e.g. the reference's own `Filter` - is used as given;
    None leaves the net without encoder (feature maps are then assigned to `im_feat_list` directly)."""

    def __init__(self, opt, projection_mode="orthogonal", criteria={"occ": nn.MSELoss()}, image_filter="auto"):  # noqa: B006 (the reference's default, `PIFuNetwNML.py:19-23`)
        super().__init__(projection_mode=projection_mode, criteria=criteria)
        self.name = "hg_pifu"
        self.opt = opt
        if isinstance(image_filter, str):
            if image_filter != "auto":
                raise ValueError("image_filter must be 'auto', None or a module")
            image_filter = encoders.build_encoder(opt, encoders.input_channels(opt), getattr(opt, "hg_down", "ave_pool"))
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
        if self.image_filter is not None:                           # `PIFuNetwNML.py:63-69`
            from .networks import define_G
            if getattr(opt, "use_front_normal", False):
                self.netF = define_G(3, 3, 64, "global", 4, 9, 1, 3, "instance")
            if getattr(opt, "use_back_normal", False):
                self.netB = define_G(3, 3, 64, "global", 4, 9, 1, 3, "instance")
        self.nmlF = None
        self.nmlB = None
        self.precise_normals = True        # finite differences in split precision (see PIFuMRNet)

    # ------------------------------------------------------------------ encoder (PyTorch, once per image)
    def filter(self, images):
        """`PIFuNetwNML.py:73-97`: optional front/back normal nets, concat, encoder."""
        if self.image_filter is None:
            raise RuntimeError("no image_filter was given to PIFuNetwNML; assign im_feat_list directly "
                               "or pass the reference's Filter module")
        extra = []
        with torch.no_grad():
            if self.netF is not None:
                self.nmlF = self._run_normal_net("netF", self.netF, images)
                extra.append(self.nmlF)
            if self.netB is not None:
                self.nmlB = self._run_normal_net("netB", self.netB, images)
                extra.append(self.nmlB)
        if extra:
            nmls = torch.cat(extra, 1)
            if nmls.shape[2:] != images.shape[2:]:
                nmls = F.interpolate(nmls, size=images.shape[2:], mode="bilinear", align_corners=True)
            images = torch.cat([images, nmls], 1)
        self.im_feat_list, self.normx = self._run_encoder("image_filter", self.image_filter, images, not self.training)

    # ------------------------------------------------------------------ fused query
    def _engine_for(self, points):
        eng = get_engine(points.device)
        eng.set_options(self.is_perspective, self.opt.loadSize, self.opt.z_size)
        eng.sync_mlp(0, self.mlp)
        return eng

    def query(self, points, calibs, transforms=None, labels=None, update_pred=True, update_phi=True):
        """`PIFuNetwNML.py:99-141`.  points [B, 3, N], calibs [B, 4, 4] (or [B, 3, 4])."""
        if transforms is not None:
            _not_hot_path("screen-space `transforms`")
        if labels is not None:
            _not_hot_path("training supervision (`labels`)")
        if len(self.im_feat_list) != 1:
            _not_hot_path("train-mode query over %d intermediate feature maps" % len(self.im_feat_list))
        check_batch_statistics(self.mlp, points.shape[0])
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
            preds.append(eng.query(1, pall[b], calibs[b], calibs[b], no_mask=True,
                                   precise=self.precise_normals)[0][None, None])
        pred = torch.cat(preds, 0).view(points.shape[0], 1, -1, 4)
        d = [pred[:, :, :, a + 1] - pred[:, :, :, 0] for a in range(3)]
        self.nml = F.normalize(-torch.cat(d, 1), dim=1, eps=1e-8)

    def get_im_feat(self):
        return self.im_feat_list[-1]

    def loadFromPIFu(self, net):
        _not_hot_path("checkpoint surgery (loadFromPIFu)")

    def forward(self, *a, **k):
        _not_hot_path("the training forward")
