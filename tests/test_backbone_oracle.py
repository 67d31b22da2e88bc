"""CPU tests of the image backbone / neck widening: the oracle restatements against torchvision's
ResNet-50 (the architecture mmdet's `ResNet(style='pytorch')` implements and the reference's config initialises from
`torchvision://resnet50`), against the UNMODIFIED reference CustomFPN (when /root/reference is present) and against the
committed fixture; the plugin modules' registry names, constructor kwargs and state_dict keys."""
import os

import numpy as np
import pytest
import torch

from oracle import dense_oracle as DO
from oracle import make_golden_backbone as MG
from oracle import ref_loader

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'backbone.npz')


def _tv_resnet50():
    torchvision = pytest.importorskip('torchvision')
    return torchvision.models.resnet50(weights=None).eval()


def test_oracle_resnet_matches_torchvision_and_fixture():
    net = _tv_resnet50()
    sd = MG.backbone_state_dict(net)
    net.load_state_dict(sd)
    img = DO.seeded_tensor(MG.IMG_SHAPE, MG.SEEDS['image'])
    with torch.no_grad():
        x = net.maxpool(net.relu(net.bn1(net.conv1(img))))
        c4 = net.layer3(net.layer2(net.layer1(x)))
        c5 = net.layer4(c4)
        o4, o5 = DO.image_resnet_forward(sd, img, depth=50, out_indices=(2, 3))
    assert torch.allclose(o4, c4, atol=1e-5, rtol=1e-5) and torch.allclose(o5, c5, atol=1e-5, rtol=1e-5)
    gold = np.load(GOLD)
    assert np.allclose(o4.numpy(), gold['c4'], atol=2e-5, rtol=1e-5) and np.allclose(o5.numpy(), gold['c5'], atol=2e-5, rtol=1e-5)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_oracle_custom_fpn_matches_the_unmodified_reference():
    ref = ref_loader.load_reference()
    neck = ref.CustomFPN(in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0, out_ids=[0]).eval()
    sd = DO.seeded_state_dict(neck, MG.SEEDS['neck'])
    neck.load_state_dict(sd)
    c4, c5 = DO.seeded_tensor((2, 1024, 6, 10), 5), DO.seeded_tensor((2, 2048, 3, 5), 6)
    with torch.no_grad():
        want = neck([c4, c5])[0]
        got = DO.custom_fpn_forward(sd, [c4, c5], out_ids=(0,), start_level=0)[0]
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    # an odd finer level (nearest interpolation to an arbitrary size, fpn.py:173-176)
    c4b = DO.seeded_tensor((1, 1024, 7, 11), 7)
    c5b = DO.seeded_tensor((1, 2048, 4, 6), 8)
    with torch.no_grad():
        assert torch.allclose(DO.custom_fpn_forward(sd, [c4b, c5b])[0], neck([c4b, c5b])[0], atol=1e-5, rtol=1e-5)


def test_oracle_fpn_matches_fixture():
    gold = np.load(GOLD)
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.necks.fpn import CustomFPN
    neck = CustomFPN(in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0, out_ids=[0])
    sd = DO.seeded_state_dict(neck, MG.SEEDS['neck'])
    out = DO.custom_fpn_forward(sd, [torch.from_numpy(gold['c4']), torch.from_numpy(gold['c5'])])[0]
    assert np.allclose(out.numpy(), gold['fpn'], atol=2e-5, rtol=1e-5)


def test_plugin_backbone_and_neck_names_and_registry():
    """`type='ResNet'` / `type='CustomFPN'` of the reference configs resolve, take the reference's kwargs and expose
    torchvision's / the reference's state_dict keys (checkpoint compatibility)."""
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200.compat import BACKBONES, NECKS
    net = BACKBONES.build(dict(type='ResNet', depth=50, num_stages=4, out_indices=(2, 3), frozen_stages=-1,
                               norm_cfg=dict(type='BN', requires_grad=True), norm_eval=False, with_cp=True, style='pytorch',
                               pretrained='torchvision://resnet50'))
    tv = _tv_resnet50()
    want = {k for k in tv.state_dict() if not k.startswith('fc.')}
    assert set(net.state_dict().keys()) == want
    for k, v in net.state_dict().items():
        assert v.shape == tv.state_dict()[k].shape, k
    assert len(BACKBONES.build(dict(type='ResNet', depth=101)).layer3) == 23
    neck = NECKS.build(dict(type='CustomFPN', in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0,
                            out_ids=[0]))
    assert set(neck.state_dict().keys()) == {'lateral_convs.0.conv.weight', 'lateral_convs.0.conv.bias',
                                             'lateral_convs.1.conv.weight', 'lateral_convs.1.conv.bias',
                                             'fpn_convs.0.conv.weight', 'fpn_convs.0.conv.bias'}
    with pytest.raises(RuntimeError):
        net.eval()(torch.zeros(1, 3, 64, 64))             # CPU tensors: no fallback


def test_oracle_resnet_training_mode_and_gradients_match_torchvision():
    """The reference of the image-backbone TRAINING parity tests (tests/test_backbone_gpu.py) is torch autograd over
    `image_resnet_forward` with BN_TRAIN (batch statistics: mmdet's train() with norm_eval=False).  Pinned here: outputs
    and every parameter gradient equal torchvision's ResNet-50 in train() mode on the same weights and image, and the
    reference CustomFPN's gradients likewise (when /root/reference is present)."""
    torchvision = pytest.importorskip('torchvision')
    net = torchvision.models.resnet50(weights=None)
    sd0 = MG.backbone_state_dict(net)
    net.load_state_dict(sd0)
    net.train()
    img = DO.seeded_tensor((2, 3, 64, 96), 31)
    x = net.maxpool(net.relu(net.bn1(net.conv1(img))))
    c4 = net.layer3(net.layer2(net.layer1(x)))
    c5 = net.layer4(c4)
    (0.5 * (c4 * c4).sum() + 0.5 * (c5 * c5).sum()).backward()
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in sd0.items()}
    DO.BN_TRAIN = True
    try:
        o4, o5 = DO.image_resnet_forward(sd, img, depth=50, out_indices=(2, 3))
    finally:
        DO.BN_TRAIN = False
    assert torch.allclose(o4, c4, atol=1e-4, rtol=1e-4) and torch.allclose(o5, c5, atol=1e-4, rtol=1e-4)
    (0.5 * (o4 * o4).sum() + 0.5 * (o5 * o5).sum()).backward()
    worst = 0.0
    for name, p in net.named_parameters():
        if name.startswith('fc.'):
            continue
        g = sd[name].grad
        assert g is not None, name
        worst = max(worst, float((g - p.grad).norm() / p.grad.norm().clamp_min(1e-12)))
    assert worst < 1e-3, worst                            # fp32 summation order only


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_oracle_custom_fpn_gradients_match_the_unmodified_reference():
    ref = ref_loader.load_reference()
    neck = ref.CustomFPN(in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0, out_ids=[0])
    sd0 = DO.seeded_state_dict(neck, MG.SEEDS['neck'])
    neck.load_state_dict(sd0)
    c4 = DO.seeded_tensor((2, 1024, 6, 10), 5).requires_grad_()
    c5 = DO.seeded_tensor((2, 2048, 3, 5), 6).requires_grad_()
    out = neck([c4, c5])[0]
    (0.5 * (out * out).sum()).backward()
    sd = {k: v.clone().requires_grad_() for k, v in sd0.items()}
    a4, a5 = c4.detach().clone().requires_grad_(), c5.detach().clone().requires_grad_()
    got = DO.custom_fpn_forward(sd, [a4, a5], out_ids=(0,), start_level=0)[0]
    (0.5 * (got * got).sum()).backward()
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-12))
    for name, p in neck.named_parameters():
        assert rel(sd[name].grad, p.grad) < 1e-4, name
    assert rel(a4.grad, c4.grad) < 1e-4 and rel(a5.grad, c5.grad) < 1e-4
