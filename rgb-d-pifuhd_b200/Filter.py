"""Stacked-hourglass image encoder, drop-in for the reference's `Filter.py`.

north_star keeps the encoders in PyTorch (they run once per image and their outputs are the hot
path's inputs); this module only has to (i) own the same parameters under the same state_dict
keys, so the reference's checkpoints load unchanged (`reconstruction.py:285-292`), and (ii)
compute the same function (`Filter.py:132-228`, checked against the reference on seeded weights,
tests/test_encoders_cpu.py).  How it is *run* on a B200 - channels_last, TF32 / bf16 autocast,
CUDA-graph replay - lives in encoders.py.

Structure (reference file:line):
  PreActBlock  `Filter.py:23-69`  three pre-activation 3x3 convs (C -> C/2 -> C/4 -> C/4), outputs
                                   concatenated, plus a (norm, ReLU, 1x1 conv) shortcut when C changes
  HourGlass    `Filter.py:71-130` depth-d recursion: skip branch at full size, the rest at half size
                                   (avg-pool down, bicubic align_corners up)
  Filter       `Filter.py:132-228` stem (7x7 stride-2 conv, block, optional avg-pool) + n_stack
                                   hourglasses with intermediate supervision taps `l{i}` and the
                                   `bl{i}` / `al{i}` re-injection between stacks
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


def _fused_ok(norm, x):
    """Eval-mode BatchNorm2d + ReLU on a contiguous fp32 CUDA tensor, no autograd: one pass in libpifu_b200.so
    (csrc/encoder_ops.cu) instead of two kernels and four passes.  PIFU_FUSED_BN_RELU=0 switches it off."""
    return (isinstance(norm, nn.BatchNorm2d) and not norm.training and norm.track_running_stats and x.is_cuda
            and x.dtype == torch.float32 and x.is_contiguous() and not torch.is_grad_enabled()
            and not torch.is_autocast_enabled() and os.environ.get("PIFU_FUSED_BN_RELU", "1") != "0")


def norm_relu(norm, x):
    """relu(norm(x)) - the pre-activation of every convolution of the hourglass (`Filter.py:61-63,187,211`)."""
    if not _fused_ok(norm, x):
        return F.relu(norm(x))
    import ctypes
    from . import _lib
    lib = _lib.load()
    y = torch.empty_like(x)
    n, c = x.shape[0], x.shape[1]
    hw = x.numel() // max(n * c, 1)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None      # noqa: E731
    with torch.cuda.device(x.device):
        _lib.check(lib.pifu_bn_relu_f32(ptr(x), ptr(norm.running_mean), ptr(norm.running_var), ptr(norm.weight), ptr(norm.bias),
                                        float(norm.eps), 1, ptr(y), n, c, hw,
                                        ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return y


def cat3_add(parts, shortcut):
    """cat(parts, 1) + shortcut - the tail of a block (`Filter.py:65-67`); one pass in libpifu_b200.so when the four
    tensors are contiguous fp32 CUDA tensors outside autograd / autocast, else the two PyTorch operators."""
    ok = (shortcut.is_cuda and not torch.is_grad_enabled() and not torch.is_autocast_enabled()
          and os.environ.get("PIFU_FUSED_BN_RELU", "1") != "0"
          and all(t.dtype == torch.float32 and t.is_contiguous() and t.data_ptr() % 16 == 0 for t in (*parts, shortcut))
          and all((t.numel() // max(t.shape[0], 1)) % 4 == 0 for t in parts) and shortcut.shape[0] <= 65535)
    if not ok:
        return torch.cat(parts, 1) + shortcut
    import ctypes
    from . import _lib
    lib = _lib.load()
    out = torch.empty_like(shortcut)
    n = shortcut.shape[0]
    sizes = [t.numel() // max(n, 1) for t in parts]
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())      # noqa: E731
    with torch.cuda.device(shortcut.device):
        _lib.check(lib.pifu_cat3_add_f32(ptr(parts[0]), ptr(parts[1]), ptr(parts[2]), ptr(shortcut), ptr(out), n, *sizes,
                                         ctypes.c_void_p(torch.cuda.current_stream(shortcut.device).cuda_stream)))
    return out


def _norm(kind, channels):
    if kind == "batch":
        return nn.BatchNorm2d(channels)
    if kind == "group":
        return nn.GroupNorm(32, channels)
    raise ValueError("norm must be 'batch' or 'group', got %r" % (kind,))


class ConvBlock(nn.Module):
    """Pre-activation residual block whose three conv outputs (C/2, C/4, C/4) are concatenated.
    Parameter names follow the reference (`conv1..3`, `bn1..4`, `downsample.{0,2}`); `bn4` and
    `downsample.0` are one module, as there."""

    def __init__(self, in_channels, out_channels, norm="batch"):
        super().__init__()
        half, quarter = int(out_channels / 2), int(out_channels / 4)
        widths = [(in_channels, half), (half, quarter), (quarter, quarter)]
        for n, (cin, cout) in enumerate(widths, start=1):
            setattr(self, "conv%d" % n, nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1, bias=False))
            setattr(self, "bn%d" % n, _norm(norm, cin))
        self.bn4 = _norm(norm, in_channels)
        self.downsample = None
        if in_channels != out_channels:
            self.downsample = nn.Sequential(self.bn4, nn.ReLU(True),
                                            nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, bias=False))

    def forward(self, x):
        # (downsample = [bn4, ReLU, 1x1 conv]; spelled out so the norm + ReLU pair can take the fused kernel)
        shortcut = x if self.downsample is None else self.downsample[2](norm_relu(self.bn4, x))
        parts, y = [], x
        for n in (1, 2, 3):
            y = getattr(self, "conv%d" % n)(norm_relu(getattr(self, "bn%d" % n), y))
            parts.append(y)
        return cat3_add(parts, shortcut)


class HourGlass(nn.Module):
    """`depth` nested levels: at level l the input goes through `b1_l` at full size and, pooled by
    two, through `b2_l` -> (level l-1 | `b2_plus_1` at the bottom) -> `b3_l`, upsampled back
    (bicubic, align_corners=True, `Filter.py:126`) and added."""

    def __init__(self, depth, n_features, norm="batch"):
        super().__init__()
        self.depth, self.features, self.norm = depth, n_features, norm
        for level in range(depth, 0, -1):
            self.add_module("b1_%d" % level, ConvBlock(n_features, n_features, norm))
            self.add_module("b2_%d" % level, ConvBlock(n_features, n_features, norm))
        self.add_module("b2_plus_1", ConvBlock(n_features, n_features, norm))
        for level in range(1, depth + 1):
            self.add_module("b3_%d" % level, ConvBlock(n_features, n_features, norm))

    def forward(self, x):
        skips, low = [], x
        for level in range(self.depth, 0, -1):           # way down: keep the full-size branches
            skips.append(self._modules["b1_%d" % level](low))
            low = self._modules["b2_%d" % level](F.avg_pool2d(low, 2, stride=2))
        low = self._modules["b2_plus_1"](low)
        for level in range(1, self.depth + 1):           # way up
            low = self._modules["b3_%d" % level](low)
            low = skips.pop() + F.interpolate(low, scale_factor=2, mode="bicubic", align_corners=True)
        return low


class Filter(nn.Module):
    """`Filter(n_stack, depth, in_channels, last_channels, norm, down_type, use_sigmoid)`;
    `forward(x) -> (outputs, normx)`: one `[B, last_channels, H', W']` map per stack and the
    128-channel stem activation.  `down_type`: 'ave_pool' (coarse net: 512 -> 128) or 'no_down'
    (fine net: 1024 -> 512).  'conv64' / 'conv128' build their layers but cannot run, exactly like
    the reference (`Filter.py:192` compares the string with a list and falls through to the
    NameError) - SURVEY §8(a-13)."""

    def __init__(self, n_stack, depth, in_channels, last_channels, norm="batch", down_type="conv64", use_sigmoid=True):
        super().__init__()
        self.n_stack, self.depth = n_stack, depth
        self.in_ch, self.last_ch = in_channels, last_channels
        self.norm, self.down_type, self.use_sigmoid = norm, down_type, use_sigmoid
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3)
        self.bn1 = _norm(norm, 64)
        if down_type == "conv64":
            self.conv2 = ConvBlock(64, 64, norm)
            self.down_conv2 = nn.Conv2d(64, 128, kernel_size=3, stride=2, padding=1)
        elif down_type == "conv128":
            self.conv2 = ConvBlock(128, 128, norm)
            self.down_conv2 = nn.Conv2d(128, 128, kernel_size=3, stride=2, padding=1)
        elif down_type in ("ave_pool", "no_down"):
            self.conv2 = ConvBlock(64, 128, norm)
        self.conv3 = ConvBlock(128, 128, norm)
        self.conv4 = ConvBlock(128, 256, norm)
        for i in range(n_stack):
            self.add_module("m%d" % i, HourGlass(depth, 256, norm))
            self.add_module("top_m_%d" % i, ConvBlock(256, 256, norm))
            self.add_module("conv_last%d" % i, nn.Conv2d(256, 256, kernel_size=1, stride=1, padding=0))
            self.add_module("bn_end%d" % i, _norm(norm, 256))
            self.add_module("l%d" % i, nn.Conv2d(256, last_channels, kernel_size=1, stride=1, padding=0))
            if i < n_stack - 1:
                self.add_module("bl%d" % i, nn.Conv2d(256, 256, kernel_size=1, stride=1, padding=0))
                self.add_module("al%d" % i, nn.Conv2d(last_channels, 256, kernel_size=1, stride=1, padding=0))

    def forward(self, x):
        x = norm_relu(self.bn1, self.conv1(x))
        if self.down_type == "ave_pool":
            x = F.avg_pool2d(self.conv2(x), 2, stride=2)
        elif self.down_type == "no_down":
            x = self.conv2(x)
        else:
            raise NameError("unknown downsampling type")
        normx = x
        carry = self.conv4(self.conv3(x))
        outputs = []
        for i in range(self.n_stack):
            m = self._modules
            ll = m["top_m_%d" % i](m["m%d" % i](carry))
            ll = norm_relu(m["bn_end%d" % i], m["conv_last%d" % i](ll))
            tap = m["l%d" % i](ll)
            outputs.append(torch.tanh(tap) if self.use_sigmoid else tap)
            if i < self.n_stack - 1:
                carry = carry + m["bl%d" % i](ll) + m["al%d" % i](tap)
        return outputs, normx
