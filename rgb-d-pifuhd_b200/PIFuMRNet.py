"""Fine (multi-level) pixel-aligned implicit function, drop-in for the reference's `PIFuMRNet.py`.

`filter_global` / `filter_local` stay PyTorch (hourglass encoders executed by encoders.EncoderRunner);
`query` / `get_preds` / `calc_normal` run coarse trunk + fine MLP in libpifu_b200.so."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import encoders
from .BasePIFuNet import BasePIFuNet, _not_hot_path, check_batch_statistics
from .MLP import MLP
from .PIFuNetwNML import EncoderHost
from .engine import get_engine


class PIFuMRNet(BasePIFuNet, EncoderHost):
    """Constructor of the reference (`PIFuMRNet.py:19-57`), including its default projection
    string 'otthogonal' (which selects perspective, `BasePIFuNet.py:79`; callers pass the mode
    explicitly, `reconstruction.py:285-286`), plus ``image_filter`` for the fine encoder ('auto': the
    'no_down' hourglass of `PIFuMRNet.py:38-39`; a module; or None, see PIFuNetwNML)."""

    def __init__(self, opt, netG, projection_mode="otthogonal", criteria={"occ": nn.MSELoss()}, image_filter="auto"):  # noqa: B006 (the reference's defaults, typo included: `PIFuMRNet.py:19-24`)
        super().__init__(projection_mode=projection_mode, criteria=criteria)
        self.name = "hg_pifu"
        self.opt = opt
        if isinstance(image_filter, str):
            if image_filter != "auto":
                raise ValueError("image_filter must be 'auto', None or a module")
            image_filter = encoders.build_encoder(opt, encoders.input_channels(getattr(netG, "opt", None)), "no_down")
        self.image_filter = image_filter
        self.mlp = MLP(filter_channels=opt.mlp_dim, merge_layer=-1, res_layers=opt.mlp_res_layers,
                       norm=opt.mlp_norm, last_op=nn.Sigmoid())
        for f in self.mlp.filters:                                  # net_util.py:13-25
            nn.init.normal_(f.weight, 0.0, 0.02)
            nn.init.constant_(f.bias, 0.0)
        self.im_feat_list = []
        self.normx = None
        self.preds_interm = None
        self.preds_low = None
        self.w = None
        self.gamma = None
        self.intermdiate_pred_list = []
        self.netG = netG
        # direct query() calls fill preds_low / netG.phi like the reference; the reconstruction
        # driver switches them off because nothing on that path reads them
        self.materialize_intermediates = True
        # calc_normal differences occupancies `delta` = 0.001 apart: the fp16 rounding noise of the fast arithmetic,
        # divided by delta, would swamp the gradient, so finite differences run in split precision (engine.set_precision)
        self.precise_normals = True

    def train(self, mode=True):
        """`PIFuMRNet.py:59-69`: the coarse net stays in eval mode unless trained end to end."""
        super().train(mode)
        if not getattr(self.opt, "train_full_pifu", False):
            self.netG.eval()
        return self

    # ------------------------------------------------------------------ encoders (PyTorch, once per image)
    def filter_global(self, images):
        """`PIFuMRNet.py:71-81`."""
        if getattr(self.opt, "train_full_pifu", False):
            self.netG.filter(images)
        else:
            with torch.no_grad():
                self.netG.filter(images)

    def filter_local(self, images, rect=None):
        """`PIFuMRNet.py:83-117`: images [B1, B2, C, H, W]; normal maps of the coarse pass are
        upsampled to loadSizeBig and concatenated (cropped per `rect` when given)."""
        if self.image_filter is None:
            raise RuntimeError("no image_filter was given to PIFuMRNet; assign im_feat_list directly "
                               "or pass the reference's Filter module")
        nmls = []
        gopt = getattr(self.netG, "opt", None)
        if getattr(gopt, "use_front_normal", False) and self.netG.nmlF is not None:
            nmls.append(self.netG.nmlF)
        if getattr(gopt, "use_back_normal", False) and self.netG.nmlB is not None:
            nmls.append(self.netG.nmlB)
        if nmls:
            big = self.opt.loadSizeBig
            nm = F.interpolate(torch.cat(nmls, 1), size=(big, big), mode="bilinear", align_corners=True)
            if rect is None:
                images = torch.cat([images, nm[:, None].expand(-1, images.size(1), -1, -1, -1)], 2)
            else:
                crops = [torch.stack([nm[i, :, r[1]:r[3], r[0]:r[2]] for r in rect[i]], 0)
                         for i in range(rect.size(0))]
                images = torch.cat([images, torch.stack(crops, 0)], 2)
        self.im_feat_list, self.normx = self._run_encoder("image_filter", self.image_filter,
                                                          images.reshape(-1, *images.shape[2:]), not self.training)

    # ------------------------------------------------------------------ fused query
    def _engine_for(self, points):
        eng = get_engine(points.device)
        eng.set_options(self.is_perspective, self.netG.opt.loadSize, self.netG.opt.z_size)
        eng.sync_mlp(0, self.netG.mlp)
        eng.sync_mlp(1, self.mlp)
        return eng

    def _features(self, b1, b2, B2):
        fine = self.im_feat_list[-1]
        return self.netG.im_feat_list[-1][b1:b1 + 1], fine[b1 * B2 + b2:b1 * B2 + b2 + 1]

    def query(self, points, calib_local, calib_global=None, transforms=None, labels=None):
        """`PIFuMRNet.py:119-186`.  Either points [B1, 3, N] + calib [B1, 4, 4] (single-level call
        form, `:131-137`) or points [B1, B2, 3, N], calib_local [B1, B2, 4, 4], calib_global [B1, 4, 4]."""
        if transforms is not None:
            _not_hot_path("screen-space `transforms`")
        if labels is not None:
            _not_hot_path("training supervision (`labels`)")
        if len(self.im_feat_list) != 1 or len(self.netG.im_feat_list) != 1:
            _not_hot_path("train-mode query over intermediate feature maps")
        if calib_global is None:
            points = points[:, None]
            calib_global = calib_local
            calib_local = calib_local[:, None]
        B1, B2 = points.shape[0], points.shape[1]
        check_batch_statistics(self.mlp, B1 * B2)
        check_batch_statistics(self.netG.mlp, B1 * B2)
        eng = self._engine_for(points)
        full = self.materialize_intermediates
        cphi = self.netG.mlp.filter_channels[self.netG.mlp.merge_layer + 1]
        preds = [[None] * B2 for _ in range(B1)]
        lows = [[None] * B2 for _ in range(B1)]
        phi_last = [None] * B1
        for b1 in range(B1):
            for b2 in range(B2):
                fc, ff = self._features(b1, b2, B2)
                eng.sync_features(0, fc)
                eng.sync_features(1, ff)
                p, low, phi = eng.query(2, points[b1, b2], calib_local[b1, b2], calib_global[b1],
                                        want_low=full, want_phi=cphi if full else 0)
                preds[b1][b2], lows[b1][b2] = p, low
                phi_last[b1] = phi
        # preds [B1*B2? no: cat over crops of [B1,1,N]] -> reference stacks crops along dim 0
        self.preds = torch.cat([torch.stack([preds[b1][b2] for b1 in range(B1)], 0)[:, None]
                                for b2 in range(B2)], 0)
        self.preds_interm = torch.cat([torch.stack([preds[b1][b2] for b1 in range(B1)], 0)[None, :, None]
                                       for b2 in range(B2)], 1)
        if full:
            self.preds_low = torch.cat([torch.stack([lows[b1][b2] for b1 in range(B1)], 0)[None, :, None]
                                        for b2 in range(B2)], 1)
            self.netG.phi = torch.stack(phi_last, 0)
            low_last = torch.stack([lows[b1][B2 - 1] for b1 in range(B1)], 0)[:, None]
            self.netG.preds = low_last
            self.netG.intermediate_preds_list = [low_last]
        else:
            self.preds_low = None

    def calc_normal(self, points, calib_local, calib_global, transforms=None, labels=None,
                    delta=0.001, fd_type="forward"):
        """`PIFuMRNet.py:188-243`: forward differences of the un-masked fine occupancy, 4 queries
        per surface point through the same fused kernels.  points [B1, B2, 3, N]."""
        if transforms is not None or labels is not None:
            _not_hot_path("`transforms` / `labels`")
        B1, B2, _, N = points.shape
        eng = self._engine_for(points)
        nmls = []
        for b2 in range(B2):
            sub = points[:, b2]
            shifted = [sub] + [sub.clone() for _ in range(3)]
            for a in range(3):
                shifted[a + 1][:, a, :] += delta
            pall = torch.stack(shifted, 3).reshape(B1, 3, -1)
            out = []
            for b1 in range(B1):
                fc, ff = self._features(b1, b2, B2)
                eng.sync_features(0, fc)
                eng.sync_features(1, ff)
                out.append(eng.query(2, pall[b1], calib_local[b1, b2], calib_global[b1], no_mask=True,
                                     precise=self.precise_normals)[0])
            pred = torch.stack(out, 0).view(B1, 1, N, 4)
            d = [pred[:, :, :, a + 1] - pred[:, :, :, 0] for a in range(3)]
            nmls.append(F.normalize(-torch.cat(d, 1), dim=1, eps=1e-8))
        self.nmls = torch.stack(nmls, 1).reshape(B1 * B2, 3, N)      # `.view(-1, 3, N)` in the reference; explicit so N == 0 works

    def get_im_feat(self):
        return self.im_feat_list[-1]

    def get_error(self):
        _not_hot_path("training loss")

    def forward(self, *a, **k):
        _not_hot_path("the training forward")
