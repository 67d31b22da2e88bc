"""TEST INFRASTRUCTURE -- generates tests/golden/backbone.npz: the image backbone + neck of DHD-S
(projects/configs/DHD/DHD-S.py:44-62) on a small seeded image batch.
  * backbone: torchvision.models.resnet50 (the architecture and parameter names of mmdet 2.25.1 `ResNet(depth=50,
    style='pytorch')`, which the reference's config builds and fills from `torchvision://resnet50`; mmdet itself is not
    vendored in /root/reference) in eval mode, outputs of layer3 / layer4 (out_indices=(2, 3));
  * neck: the UNMODIFIED reference CustomFPN (projects/mmdet3d_plugin/models/necks/fpn.py) imported through
    oracle/ref_loader.py with in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0, out_ids=[0].
Run in the build container:  python -m oracle.make_golden_backbone
"""
import os

import numpy as np
import torch

from . import dense_oracle as DO
from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'backbone.npz')
IMG_SHAPE = (2, 3, 96, 160)          # -> C4 (2, 1024, 6, 10), C5 (2, 2048, 3, 5)
SEEDS = dict(backbone=61, neck=62, image=63)


def build_reference():
    import torchvision
    ref = ref_loader.load_reference()
    net = torchvision.models.resnet50(weights=None).eval()
    neck = ref.CustomFPN(in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0, out_ids=[0]).eval()
    return net, neck


def backbone_state_dict(net):
    sd = DO.seeded_state_dict(net, SEEDS['backbone'])
    return {k: v for k, v in sd.items()}


def reference_forward(net, neck, img):
    with torch.no_grad():
        x = net.maxpool(net.relu(net.bn1(net.conv1(img))))
        c2 = net.layer1(x)
        c3 = net.layer2(c2)
        c4 = net.layer3(c3)
        c5 = net.layer4(c4)
        out = neck([c4, c5])[0]
    return c4, c5, out


def main():
    net, neck = build_reference()
    net.load_state_dict(backbone_state_dict(net))
    neck.load_state_dict(DO.seeded_state_dict(neck, SEEDS['neck']))
    img = DO.seeded_tensor(IMG_SHAPE, SEEDS['image'])
    c4, c5, out = reference_forward(net, neck, img)
    np.savez_compressed(OUT, c4=c4.numpy(), c5=c5.numpy(), fpn=out.numpy())
    print('wrote', OUT, {k: tuple(v.shape) for k, v in dict(c4=c4, c5=c5, fpn=out).items()},
          'scales', float(c4.abs().max()), float(c5.abs().max()), float(out.abs().max()))


if __name__ == '__main__':
    main()
