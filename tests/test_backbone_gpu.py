"""Image backbone + neck on the tcgen05 convolution kernel (SURVEY 8(f)-4 widening) against the committed fixture
(tests/golden/backbone.npz: torchvision ResNet-50 + the unmodified reference CustomFPN, oracle/make_golden_backbone.py)
and the oracle restatement: fp32 mode (6-term split bf16) within 1e-4 of the output scale through all 53 layers, bf16
speed mode against its own bound; the helper kernels (stem im2col, MaxPool 3x3/2, nearest up-sampling + add) exactly."""
import os

import numpy as np
import pytest
import torch

from oracle import dense_oracle as DO
from oracle import make_golden_backbone as MG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'backbone.npz')


def build(precision):
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.backbones.image_resnet import ResNet
    from projects.mmdet3d_plugin.models.necks.fpn import CustomFPN
    net = ResNet(depth=50, out_indices=(2, 3), style='pytorch', precision=precision).eval()
    neck = CustomFPN(in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0, out_ids=[0],
                     precision=precision).eval()
    net.load_state_dict(DO.seeded_state_dict(net, MG.SEEDS['backbone']))
    neck.load_state_dict(DO.seeded_state_dict(neck, MG.SEEDS['neck']))
    return net.cuda(), neck.cuda()


def test_helper_kernels_match_torch(cuda_lib):
    from dhd_b200 import backbone as BB
    from dhd_b200 import dense as D
    from dhd_b200.modules import unpack
    g = torch.Generator().manual_seed(4)
    img = torch.randn(2, 3, 37, 53, generator=g).cuda()
    for parts in (1, 3):
        col = BB.stem_im2col(img, 7, 2, 3, parts)
        want = torch.nn.functional.unfold(img, 7, padding=3, stride=2)                     # (N, c*49 (c, ky, kx), L)
        want = want.view(2, 3, 49, col.H, col.W).permute(0, 3, 4, 2, 1).reshape(2, col.H, col.W, 147)   # K = (ky, kx, c)
        got = sum(col.data[..., p * col.C:p * col.C + 147].float() for p in range(parts))
        tol = 2 ** -8 if parts == 1 else 1e-6
        assert float((got - want).abs().max()) <= tol * float(want.abs().max())
        assert float(col.data[..., 147:col.C].abs().max()) == 0.0
    x = torch.randn(2, 64, 19, 27, generator=g).cuda()
    for parts in (1, 3):
        a = D.pack_input(x, parts)
        want = torch.nn.functional.max_pool2d(unpack(a), 3, stride=2, padding=1)
        assert torch.equal(unpack(BB.maxpool3s2(a)), want) or float((unpack(BB.maxpool3s2(a)) - want).abs().max()) < 1e-6
    lo, hi = torch.randn(2, 64, 4, 6, generator=g).cuda(), torch.randn(2, 64, 7, 11, generator=g).cuda()
    for parts in (1, 3):
        a, b = D.pack_input(lo, parts), D.pack_input(hi, parts)
        want = unpack(b) + torch.nn.functional.interpolate(unpack(a), size=(7, 11), mode='nearest')
        BB.upsample_nearest_add(a, b)
        tol = 2 ** -7 if parts == 1 else 1e-6
        assert float((unpack(b) - want).abs().max()) <= tol * float(want.abs().max())


def test_resnet50_and_fpn_fp32_mode_match_reference_fixture(cuda_lib):
    gold = np.load(GOLD)
    net, neck = build('fp32')
    img = DO.seeded_tensor(MG.IMG_SHAPE, MG.SEEDS['image']).cuda()
    c4, c5 = net(img)
    for got, name in ((c4, 'c4'), (c5, 'c5')):
        ref = torch.from_numpy(gold[name])
        err = float((got.cpu() - ref).abs().max())
        assert err <= 1e-4 * float(ref.abs().max()), (name, err, float(ref.abs().max()))
    out = neck([c4, c5])[0]
    ref = torch.from_numpy(gold['fpn'])
    assert float((out.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    # Acts straight through (no NCHW round trip) give the same map
    out2 = neck(net(img, return_act=True))[0]
    assert float((out2 - out).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_resnet50_and_fpn_bf16_speed_mode(cuda_lib):
    gold = np.load(GOLD)
    net, neck = build('bf16')
    img = DO.seeded_tensor(MG.IMG_SHAPE, MG.SEEDS['image']).cuda()
    out = neck(net(img, return_act=True))[0].cpu()
    ref = torch.from_numpy(gold['fpn'])
    rel = float((out - ref).norm() / ref.norm())
    assert rel <= 3e-2, rel                                   # 53 bf16 layers: relative L2 error, stated
    assert float((out - ref).abs().max()) <= 0.15 * float(ref.abs().max())


def test_odd_image_size_and_resnet101_against_the_oracle(cuda_lib):
    """Odd feature-map sizes (stride-2 layers round up, the nearest up-sampling meets a non-2x ratio) and depth 101."""
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.backbones.image_resnet import ResNet
    net, neck = build('fp32')
    img = DO.seeded_tensor((1, 3, 104, 200), 9).cuda()
    with torch.no_grad():
        want = DO.image_resnet_forward({k: v.cpu() for k, v in net.state_dict().items()}, img.cpu(), 50, (2, 3))
        fw = DO.custom_fpn_forward({k: v.cpu() for k, v in neck.state_dict().items()}, want)[0]
    got = net(img)
    for a, b in zip(got, want):
        assert a.shape == b.shape and float((a.cpu() - b).abs().max()) <= 1e-4 * float(b.abs().max())
    assert float((neck(got)[0].cpu() - fw).abs().max()) <= 1e-4 * float(fw.abs().max())
    r101 = ResNet(depth=101, out_indices=(3,), precision='fp32').eval()
    r101.load_state_dict(DO.seeded_state_dict(r101, 71))
    img = DO.seeded_tensor((1, 3, 64, 96), 10)
    with torch.no_grad():
        want = DO.image_resnet_forward(r101.state_dict(), img, 101, (3,))[0]
    got = r101.cuda()(img.cuda())[0].cpu()
    assert float((got - want).abs().max()) <= 2e-4 * float(want.abs().max())


def test_dhd_from_images_equals_dhd_from_features(cuda_lib):
    """The DHD-S detector with the reference config's `img_backbone` / `img_neck` entries (DHD-S.py:44-62) owns its image
    encoder: simple_test on images == simple_test on the features that backbone + neck produce (the injection point of
    an external backbone stays)."""
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200 import compat as C
    from dhd_b200 import synth as O
    from tests.test_encoders_gpu import _dhd_s_model_cfg
    cfg = _dhd_s_model_cfg()
    cfg['img_backbone'] = dict(type='ResNet', depth=50, num_stages=4, out_indices=(2, 3), frozen_stages=-1,
                               norm_cfg=dict(type='BN', requires_grad=True), norm_eval=False, with_cp=True, style='pytorch',
                               pretrained='torchvision://resnet50')
    cfg['img_neck'] = dict(type='CustomFPN', in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0,
                           out_ids=[0])
    model = C.DETECTORS.build(cfg).eval()
    model.load_state_dict(DO.seeded_state_dict(model, 77))
    model = model.cuda()
    assert type(model.img_backbone).__name__ == 'ResNet' and type(model.img_neck).__name__ == 'CustomFPN'
    B, N = 1, 6
    rig = [t.cuda() for t in O.synthetic_rig(B, N, (256, 704), seed=5)]
    imgs = DO.seeded_tensor((B, N, 3, 256, 704), 12).cuda()
    with torch.no_grad():
        feats = model.image_encoder(imgs)[0]
        assert feats.shape == (B, N, 256, 16, 44) and torch.isfinite(feats).all()
        a = model.simple_test(None, None, img=[imgs] + rig)
        b = model.simple_test(None, None, img=[feats] + rig)
    assert np.array_equal(np.asarray(a[0]), np.asarray(b[0]))
