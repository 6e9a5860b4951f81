"""Tuned execution of the PyTorch encoders on the device: channels_last + CUDA-graph
replay must reproduce the plain eager fp32 forward of the same module (the reference's execution on a GPU) to
float32 round-off; TF32 / bf16 are opt-in and only checked for sanity.  Full-size encoders of the reference
configuration (`options.py` defaults via `train.py:102-120`): coarse 4-stack hourglass on 512^2, fine 1-stack
'no_down' on 1024^2."""
import pytest
import torch

from pifu_b200 import PIFuMRNet, PIFuNetwNML, config, encoders, synthetic as syn
from pifu_b200.Filter import Filter

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.fixture(scope="module")
def nets():
    torch.set_grad_enabled(False)
    coarse = Filter(4, 2, 6, 256, "group", "ave_pool", False)
    fine = Filter(1, 2, 6, 16, "group", "no_down", False)
    syn.fill_state(coarse, 3)
    syn.fill_state(fine, 4)
    return coarse.cuda().eval(), fine.cuda().eval()


@pytest.mark.parametrize("which,size", [("coarse", 512), ("fine", 1024)])
def test_graph_channels_last_equals_eager(nets, which, size):
    net = nets[0] if which == "coarse" else nets[1]
    x1 = syn.encoder_input((1, 6, size, size), 21).cuda()
    x2 = syn.encoder_input((1, 6, size, size), 22).cuda()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref1, n1 = net(x1)
    ref2, n2 = net(x2)
    run = encoders.EncoderRunner(net, channels_last=True, precision="fp32", graph=True)
    a1, an1 = run(x1, last_only=True)
    a2, an2 = run(x2, last_only=True)              # replay of the captured graph with new input
    a1b, _ = run(x1, last_only=True)
    assert len(a1) == 1 and a1[0].shape == ref1[-1].shape and a1[0].dtype == torch.float32
    assert rel(a1[0], ref1[-1]) < 1e-4 and rel(a2[0], ref2[-1]) < 1e-4 and rel(an2, n2) < 1e-4
    assert torch.equal(a1[0], a1b[0])              # replays are deterministic and do not alias their outputs
    assert rel(ref1[-1], ref2[-1]) > 1e-2          # (the two inputs really differ)
    for prec, tol in (("tf32", 2e-2), ("bf16", 2e-1)):
        r = encoders.EncoderRunner(net, channels_last=True, precision=prec, graph=False)
        b, _ = r(x1, last_only=True)
        assert rel(b[0], ref1[-1]) < tol
    assert torch.backends.cudnn.allow_tf32 is False   # the runner restores the global switches


def test_filter_calls_use_the_runner_and_feed_the_query_path():
    """filter_global / filter_local -> im_feat_list -> query: the whole per-image flow on the device."""
    torch.set_grad_enabled(False)
    og = config.coarse_opt(use_front_normal=True, norm="group")
    netG = PIFuNetwNML(og, "orthogonal")
    netG.netF = None                                   # RGB-D: depth rides in the normal channels
    netMR = PIFuMRNet(config.fine_opt(norm="group"), netG, "orthogonal")
    netMR.cuda().eval()
    img512 = syn.encoder_input((1, 6, 512, 512), 5).cuda()
    img1024 = syn.encoder_input((1, 1, 6, 1024, 1024), 6).cuda()
    netMR.filter_global(img512)
    netMR.filter_local(img1024)
    assert netG.im_feat_list[0].shape == (1, 256, 128, 128) and netMR.im_feat_list[0].shape == (1, 16, 512, 512)
    assert "image_filter" in netG._enc_runners and "image_filter" in netMR._enc_runners
    f0 = netG.im_feat_list[0].clone()
    netMR.filter_global(img512)                        # second image: graph replay
    assert torch.equal(f0, netG.im_feat_list[0])
    pts = syn.random_points(1000).cuda()
    netMR.query(pts, syn.default_calib().cuda())
    p = netMR.get_preds()
    assert p.shape == (1, 1, 1000) and bool(torch.isfinite(p).all())


@pytest.mark.parametrize("shape", [(2, 64, 33, 37), (1, 256, 64, 64), (3, 5, 1, 8)])
def test_fused_bn_relu_matches_torch(shape):
    """csrc/encoder_ops.cu against F.relu(BatchNorm2d.eval()(x)): same formula, one pass."""
    from pifu_b200.Filter import norm_relu, _fused_ok
    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(1)
    bn = torch.nn.BatchNorm2d(shape[1])
    bn.running_mean.copy_(torch.randn(shape[1], generator=g))
    bn.running_var.copy_(torch.rand(shape[1], generator=g) + 0.3)
    bn.weight.copy_(1 + 0.2 * torch.randn(shape[1], generator=g))
    bn.bias.copy_(0.3 * torch.randn(shape[1], generator=g))
    bn = bn.cuda().eval()
    x = (2 * torch.randn(shape, generator=g)).cuda()
    assert _fused_ok(bn, x)
    y = norm_relu(bn, x)
    ref = torch.nn.functional.relu(bn(x))
    assert y.shape == ref.shape and float((y - ref).abs().max()) < 1e-5          # a few ulps: cuDNN folds the affine map differently
    assert not _fused_ok(bn.train(), x) and not _fused_ok(bn.eval(), x.half())
    assert not _fused_ok(torch.nn.GroupNorm(1, shape[1]).cuda(), x)


def test_fused_cat3_add_matches_torch():
    from pifu_b200.Filter import cat3_add
    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(2)
    for n, c, h, w in ((1, 256, 32, 32), (3, 64, 6, 10)):
        parts = [torch.randn(n, c // d, h, w, generator=g).cuda() for d in (2, 4, 4)]
        sc = torch.randn(n, c, h, w, generator=g).cuda()
        assert torch.equal(cat3_add(parts, sc), torch.cat(parts, 1) + sc)
    odd = [torch.randn(1, k, 3, 3, generator=g).cuda() for k in (2, 1, 1)]          # 9 floats per channel: falls back
    sc = torch.randn(1, 4, 3, 3, generator=g).cuda()
    assert torch.equal(cat3_add(odd, sc), torch.cat(odd, 1) + sc)
