"""The PyTorch encoders of this package (hourglass `Filter`, pix2pixHD normal
generator) against outputs of the UNMODIFIED reference modules (`Filter.py:132-228`,
`networks.py:35-60,131-166`) stored by oracle/make_golden_encoders.py: same state_dict keys and shapes
(checkpoints interchange) and the same function on key-addressed seeded parameters.  CPU, fp32."""
import os

import numpy as np
import pytest
import torch

from pifu_b200 import PIFuMRNet, PIFuNetwNML, config, encoders, synthetic as syn
from pifu_b200.Filter import Filter
from pifu_b200.networks import define_G

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "encoders.npz"))
TOL = 2e-5          # fp32, same operators in the same order; outputs are O(1)


def build(kind, args):
    return Filter(*args) if kind == "filter" else define_G(*args)


@pytest.mark.parametrize("name", sorted(syn.ENCODER_CASES))
def test_matches_reference(name):
    kind, args, shape = syn.ENCODER_CASES[name]
    torch.set_grad_enabled(False)
    torch.set_num_threads(1)
    net = build(kind, args)
    assert bytes(GOLD[name + "_sig"]).hex() == syn.state_signature(net), "state_dict keys / shapes differ from the reference's"
    seed = int(GOLD[name + "_seed"])
    syn.fill_state(net, seed)
    net.eval()
    y = net(syn.encoder_input(shape, seed + 100))
    if kind == "filter":
        feats, normx = y
        assert len(feats) == args[0]
        for i, f in enumerate(feats):
            assert np.abs(f.numpy() - GOLD["%s_out%d" % (name, i)]).max() < TOL
        assert np.abs(normx.numpy() - GOLD[name + "_normx"]).max() < TOL
    else:
        assert np.abs(y.numpy() - GOLD[name + "_out0"]).max() < TOL


def test_unrunnable_down_types_like_reference():
    """`Filter.py:192` compares a str with a list: 'conv64' / 'conv128' build but cannot run."""
    for down in ("conv64", "conv128"):
        net = Filter(1, 1, 3, 8, "batch", down, False)
        assert hasattr(net, "down_conv2")
        with pytest.raises(NameError):
            net(torch.zeros(1, 3, 32, 32))


def test_define_g_other_kinds_out_of_scope():
    with pytest.raises(NotImplementedError):
        define_G(3, 3, 8, "local")


def test_nets_build_their_encoders():
    """Constructor parity (`PIFuNetwNML.py:31-41,63-69`, `PIFuMRNet.py:27-39`): input channels follow the
    normal-map switches, the fine encoder is 'no_down', eval mode keeps the last stack only."""
    torch.set_grad_enabled(False)
    og = config.coarse_opt(num_stack=2, hg_depth=1, use_front_normal=True)
    netG = PIFuNetwNML(og, "orthogonal")
    assert netG.image_filter.conv1.in_channels == 6 and netG.image_filter.down_type == "ave_pool"
    assert netG.netF is not None and netG.netB is None
    netG.netF = None                                  # the RGB-D route of SURVEY §8(c): depth rides in the normal channels
    netMR = PIFuMRNet(config.fine_opt(hg_depth=1), netG, "orthogonal")
    assert netMR.image_filter.conv1.in_channels == 6 and netMR.image_filter.down_type == "no_down"
    netMR.eval()
    img = torch.randn(1, 6, 64, 64)
    netMR.filter_global(img)
    assert len(netG.im_feat_list) == 1 and netG.im_feat_list[0].shape == (1, 256, 16, 16)
    assert netG.normx.shape == (1, 128, 16, 16)
    netMR.filter_local(torch.randn(1, 1, 6, 64, 64))
    assert len(netMR.im_feat_list) == 1 and netMR.im_feat_list[0].shape == (1, 16, 32, 32)
    netG.train()
    netG.filter(img)
    assert len(netG.im_feat_list) == 2                # train mode keeps every stack (`PIFuNetwNML.py:96-97`)
    # query-only namespaces (no encoder fields) and explicit None leave the net without encoder
    assert PIFuNetwNML(config.coarse_opt(), "orthogonal", image_filter=None).image_filter is None
    sd = netMR.state_dict()
    assert "image_filter.conv1.weight" in sd and "netG.image_filter.m0.b2_plus_1.conv1.weight" in sd and "mlp.filters.0.weight" in sd


def test_runner_on_cpu_is_the_plain_module():
    torch.set_grad_enabled(False)
    net = Filter(1, 1, 3, 8, "group", "no_down", False).eval()
    x = torch.randn(1, 3, 32, 32)
    a, na = net(x)
    b, nb = encoders.EncoderRunner(net, precision="bf16", graph=True)(x)       # CPU: no autocast, no graph
    assert torch.equal(a[0], b[0]) and torch.equal(na, nb)


def test_index_matches_oracle_closed_form():
    """`BasePIFuNet.index` (vertex colours of gen_mesh_imgColor, `reconstruction.py:110-116`) against the oracle's
    closed form of aten's bilinear grid_sample (align_corners=True, zeros padding)."""
    from oracle import pifu_oracle as orc
    from pifu_b200.BasePIFuNet import index
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(1, 3, 37, 53, generator=g)
    uv = torch.rand(1, 2, 500, generator=g) * 2.4 - 1.2          # some samples outside the image
    assert (index(feat, uv) - orc.index_closed_form(feat, uv)).abs().max() < 1e-5
