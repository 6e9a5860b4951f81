"""How the image encoders run on a B200.

north_star keeps `Filter` / the normal nets in PyTorch: they run once per image.  Once the query path
takes 14 ms per 512^3 mesh they are the largest per-frame cost (BASELINE configs[3], frames/s), so
this module tunes how they are *executed*, not what they compute:

* `channels_last`  NHWC activations and weights (off by default: measured on B200, the hourglass is
                   bound by its element-wise traffic - norms, ReLUs, concatenations, bicubic upsampling -
                   and those kernels are slower in NHWC: 31.2 ms against 18.2 ms per frame);
* `precision`      'tf32' (cuDNN TF32 convolutions: fp32 storage, 10-bit-mantissa products, fp32
                   accumulate - what stock PyTorch runs for the reference's convolutions on this GPU,
                   `torch.backends.cudnn.allow_tf32` defaults to True), 'fp32' (IEEE) or 'bf16' (autocast);
* `graph`          the encoder forward captured once per input shape into a CUDA graph and replayed:
                   a 4-stack hourglass is ~700 small kernels whose launch overhead (~4 us each) is
                   otherwise most of its GPU time.

`build_encoder(opt, in_channels, down_type)` creates the `Filter` the reference's constructors create
(`PIFuNetwNML.py:40-41`, `PIFuMRNet.py:38-39`) with the reference's initialisation
(`net_util.py:10-32`).  `EncoderRunner(module, ...)` wraps any encoder module returning
`(feature_list, normx)`.
"""
import contextlib

import torch
import torch.nn as nn

from .Filter import Filter

PRECISIONS = ("fp32", "tf32", "bf16")


def init_like_reference(module, gain=0.02):
    """`net_util.init_weights`: Conv*/Linear weights N(0, gain), biases 0; BatchNorm2d weights N(1, gain), biases 0."""
    for m in module.modules():
        name = m.__class__.__name__
        if hasattr(m, "weight") and m.weight is not None and (name.find("Conv") != -1 or name.find("Linear") != -1):
            nn.init.normal_(m.weight.data, 0.0, gain)
            if getattr(m, "bias", None) is not None:
                nn.init.constant_(m.bias.data, 0.0)
        elif name.find("BatchNorm2d") != -1:
            nn.init.normal_(m.weight.data, 1.0, gain)
            nn.init.constant_(m.bias.data, 0.0)


def input_channels(opt):
    """3 image channels + 3 per enabled normal map (`PIFuNetwNML.py:31-38`)."""
    c = 3
    if getattr(opt, "use_front_normal", False):
        c += 3
    if getattr(opt, "use_back_normal", False):
        c += 3
    return c


def build_encoder(opt, in_channels, down_type):
    """The hourglass `Filter` of a level, or None when `opt` carries no encoder fields (a namespace made
    only for the query path)."""
    need = ("num_stack", "hg_depth", "hg_dim", "norm")
    if not all(hasattr(opt, k) for k in need):
        return None
    enc = Filter(opt.num_stack, opt.hg_depth, in_channels, opt.hg_dim, opt.norm, down_type, False)
    init_like_reference(enc)
    return enc


class EncoderRunner:
    """Runs an encoder module `(images) -> (feature_list, normx)` with the chosen layout / precision and,
    optionally, CUDA-graph replay.  Results are float32 NCHW-shaped tensors (channels_last strides are
    fine for the consumers: the engine re-lays the feature map out itself)."""

    def __init__(self, module, channels_last=False, precision="tf32", graph=True, autotune=True):
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % (PRECISIONS,))
        self.module, self.channels_last, self.precision, self.graph = module, channels_last, precision, graph
        self.autotune = autotune
        self._graphs = {}                # (shape, dtype, device, training, keep_all) -> (graph, static_in, static_out)
        self._cl_done = False

    # -- numerics context
    @contextlib.contextmanager
    def _numerics(self, device):
        if device.type != "cuda":
            yield
            return
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
        torch.backends.cudnn.allow_tf32 = self.precision != "fp32"
        torch.backends.cuda.matmul.allow_tf32 = self.precision != "fp32"
        torch.backends.cudnn.benchmark = self.autotune        # fixed shapes, one image after another: let cuDNN pick
        try:
            if self.precision == "bf16":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    yield
            else:
                yield
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old

    def _forward(self, images, last_only):
        if self.channels_last and images.dim() == 4 and images.is_cuda:
            images = images.contiguous(memory_format=torch.channels_last)
        with self._numerics(images.device):
            feats, normx = self.module(images)
        if last_only:
            feats = [feats[-1]]
        return [f.float() for f in feats], (normx.float() if normx is not None else None)

    def __call__(self, images, last_only=False):
        """`last_only`: keep only the last stack's map (what eval mode keeps, `PIFuNetwNML.py:96-97`)."""
        if self.channels_last and not self._cl_done and images.is_cuda:
            self.module.to(memory_format=torch.channels_last)
            self._cl_done = True
        use_graph = (self.graph and images.is_cuda and not torch.is_grad_enabled() and not self.module.training)
        if not use_graph:
            return self._forward(images, last_only)
        key = (tuple(images.shape), images.dtype, images.device, last_only)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = images.clone()
            side = torch.cuda.Stream(device=images.device)
            side.wait_stream(torch.cuda.current_stream(images.device))
            with torch.cuda.stream(side):                       # warm-up off the capture (cuDNN autotune, allocations)
                for _ in range(2):
                    self._forward(static_in, last_only)
            torch.cuda.current_stream(images.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self._forward(static_in, last_only)
            entry = (g, static_in, static_out)
            self._graphs[key] = entry
        g, static_in, static_out = entry
        static_in.copy_(images)
        g.replay()
        feats, normx = static_out
        return [f.clone() for f in feats], (normx.clone() if normx is not None else None)

    def invalidate(self):
        """Drop captured graphs (after load_state_dict the parameters are updated in place and graphs stay
        valid; call this only when parameters were re-allocated, e.g. `.to(other_device)`)."""
        self._graphs.clear()
        self._cl_done = False
