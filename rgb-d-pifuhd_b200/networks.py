"""Front / back normal-map generators of the coarse net (`netF` / `netB`, `PIFuNetwNML.py:63-69`):
the pix2pixHD "global" generator the reference builds with `define_G(3, 3, 64, 'global', 4, 9, 1, 3,
'instance')` (`networks.py:35-60,131-166`).  Inference only, PyTorch (north_star: the encoders run
once per image); same `model.<n>.*` / `conv_block.<n>.*` state_dict keys so the reference's
checkpoints load.  The GAN training parts of `networks.py` (local enhancer, discriminators, VGG /
GAN losses) are out of scope."""
import functools

import torch.nn as nn


def get_norm_layer(norm_type="instance"):
    if norm_type == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True)
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False)
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


class ResnetBlock(nn.Module):
    """x + (pad, conv3, norm, ReLU, pad, conv3, norm)(x)  (`networks.py:168-214`)."""

    _PADS = {"reflect": nn.ReflectionPad2d, "replicate": nn.ReplicationPad2d}

    def __init__(self, dim, padding_type, norm_layer, activation=None, use_dropout=False):
        super().__init__()
        activation = activation if activation is not None else nn.ReLU(True)
        layers = []
        for half in range(2):
            if padding_type in self._PADS:
                layers.append(self._PADS[padding_type](1))
                layers.append(nn.Conv2d(dim, dim, kernel_size=3, padding=0))
            elif padding_type == "zero":
                layers.append(nn.Conv2d(dim, dim, kernel_size=3, padding=1))
            else:
                raise NotImplementedError("padding [%s] is not implemented" % padding_type)
            layers.append(norm_layer(dim))
            if half == 0:
                layers.append(activation)
                if use_dropout:
                    layers.append(nn.Dropout(0.5))
        self.conv_block = nn.Sequential(*layers)

    def forward(self, x):
        return x + self.conv_block(x)


class GlobalGenerator(nn.Module):
    """7x7 stem -> `n_downsampling` stride-2 convs -> `n_blocks` residual blocks -> mirrored
    transposed convs -> 7x7 head (+ `last_op`)."""

    def __init__(self, input_nc, output_nc, ngf=64, n_downsampling=3, n_blocks=9, norm_layer=nn.BatchNorm2d,
                 padding_type="reflect", last_op=nn.Tanh()):
        assert n_blocks >= 0
        super().__init__()
        act = nn.ReLU(True)
        seq = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0), norm_layer(ngf), act]
        ch = ngf
        for _ in range(n_downsampling):
            seq += [nn.Conv2d(ch, ch * 2, kernel_size=3, stride=2, padding=1), norm_layer(ch * 2), act]
            ch *= 2
        seq += [ResnetBlock(ch, padding_type=padding_type, activation=act, norm_layer=norm_layer) for _ in range(n_blocks)]
        for _ in range(n_downsampling):
            seq += [nn.ConvTranspose2d(ch, ch // 2, kernel_size=3, stride=2, padding=1, output_padding=1),
                    norm_layer(ch // 2), act]
            ch //= 2
        seq += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0)]
        if last_op is not None:
            seq.append(last_op)
        self.model = nn.Sequential(*seq)

    def forward(self, input):
        return self.model(input)


def _weights_init(m):
    name = m.__class__.__name__                               # `networks.py:13-22`
    if name.find("Conv") != -1:
        m.weight.data.normal_(0.0, 0.02)
    elif name.find("BatchNorm2d") != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def define_G(input_nc, output_nc, ngf, netG, n_downsample_global=3, n_blocks_global=9, n_local_enhancers=1,
             n_blocks_local=3, norm="instance", gpu_ids=(), last_op=nn.Tanh()):
    """`networks.py:35-60`, generator kind 'global' (the only one the reconstruction path builds)."""
    if netG != "global":
        raise NotImplementedError("generator %r belongs to the normal-net training code, outside the reconstruction "
                                  "path (SURVEY.md §2 row 8)" % (netG,))
    net = GlobalGenerator(input_nc, output_nc, ngf, n_downsample_global, n_blocks_global, get_norm_layer(norm),
                          last_op=last_op)
    if len(gpu_ids) > 0:
        net.cuda(gpu_ids[0])
    net.apply(_weights_init)
    return net
