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
