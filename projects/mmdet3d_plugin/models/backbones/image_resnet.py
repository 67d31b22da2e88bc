"""Image backbone `ResNet` (reference: `img_backbone=dict(type='ResNet', depth=50, num_stages=4, out_indices=(2, 3),
frozen_stages=-1, norm_cfg=dict(type='BN', requires_grad=True), norm_eval=False, with_cp=True, style='pytorch',
pretrained='torchvision://resnet50')`, projects/configs/DHD/DHD-S.py:44-55; the class is mmdet 2.25.1
`mmdet.models.backbones.ResNet`, which is not part of the reference tree: the architecture is torchvision's ResNet, whose
parameter names -- conv1 / bn1 / layer{1..4}.{i}.conv{1,2,3} / bn{1,2,3} / downsample.{0,1} -- the `torchvision://`
checkpoints carry).  Same registry name, constructor kwargs and state_dict keys; eval-mode forward on
dhd_b200.backbone.ResNetEngine (tcgen05 convolutions); under autograd (train() mode, or eval() with trainable
parameters and gradients enabled) the forward runs dhd_b200.train_backbone.ImageResNetTrainer and `.backward()` its
hand-written backward: BatchNorm on batch statistics in train() unless `norm_eval` (mmdet's rule), frozen otherwise."""
import torch
import torch.nn as nn

from dhd_b200.compat import BACKBONES, EngineOwner


class Bottleneck(nn.Module):
    """Parameter container of the torchvision / mmdet Bottleneck (style='pytorch': stride on conv2)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample


class ResNet(EngineOwner, nn.Module):
    arch_settings = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}

    def __init__(self, depth, in_channels=3, stem_channels=None, base_channels=64, num_stages=4, strides=(1, 2, 2, 2),
                 dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3), style='pytorch', deep_stem=False, avg_down=False,
                 frozen_stages=-1, conv_cfg=None, norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, dcn=None,
                 stage_with_dcn=(False, False, False, False), plugins=None, with_cp=False, zero_init_residual=True,
                 pretrained=None, init_cfg=None, precision='bf16'):
        super().__init__()
        if depth not in self.arch_settings:
            raise NotImplementedError('ResNet depth %s: the DHD configs use the Bottleneck depths 50 / 101' % depth)
        if style != 'pytorch' or deep_stem or avg_down or dcn is not None or plugins is not None or \
                tuple(dilations[:num_stages]) != (1,) * num_stages or norm_cfg.get('type', 'BN') not in ('BN', 'BN2d'):
            raise NotImplementedError('ResNet variant outside the DHD configs (style=pytorch, plain stem, BN, no DCN)')
        self.depth, self.out_indices, self.frozen_stages = depth, tuple(out_indices), frozen_stages
        self.norm_eval, self.with_cp, self.precision = norm_eval, with_cp, precision
        stem = stem_channels or base_channels
        self.conv1 = nn.Conv2d(in_channels, stem, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(stem)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.res_layers = []
        inplanes = stem
        for i, nblocks in enumerate(self.arch_settings[depth][:num_stages]):
            planes, stride = base_channels * 2 ** i, strides[i]
            ds = None
            if stride != 1 or inplanes != planes * 4:
                ds = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False), nn.BatchNorm2d(planes * 4))
            blocks = [Bottleneck(inplanes, planes, stride, ds)]
            inplanes = planes * 4
            blocks += [Bottleneck(inplanes, planes) for _ in range(nblocks - 1)]
            name = 'layer%d' % (i + 1)
            self.add_module(name, nn.Sequential(*blocks))
            self.res_layers.append(name)
        self.feat_dim = inplanes
        self._engine = None

    def init_weights(self):
        pass

    def forward(self, x, return_act=False):
        """x: (N, 3, H, W) images -> tuple of (N, C_i, H / s_i, W / s_i) feature maps of `out_indices` (Acts with
        return_act=True)."""
        from dhd_b200.backbone import ResNetEngine
        from dhd_b200.modules import unpack
        if not x.is_cuda:
            raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
        from dhd_b200 import autograd as A
        if A.wants_grad(self):
            feats = A.image_resnet_forward(self, x)      # differentiable form (dhd_b200.train_backbone)
            if return_act:
                raise NotImplementedError('return_act=True is the inference hand-off; under autograd tensors are returned')
            return tuple(feats)
        with torch.no_grad():
            dev = x.device
            feats = self.cached_engine(dev, lambda: ResNetEngine(self, self.precision, dev))(x)
            return tuple(feats) if return_act else tuple(unpack(f) for f in feats)


# `ResNet` is mmdet's registry name: where the genuine mmdet class is registered (a full mmdet3d installation) it stays --
# this module then serves as `type='dhd_b200.ResNet'`; in the stand-alone build it answers to the reference's config.
BACKBONES.register_module(name='dhd_b200.ResNet', force=True, module=ResNet)
try:
    if BACKBONES.get('ResNet') is None:
        BACKBONES.register_module(name='ResNet', module=ResNet)
except Exception:  # noqa: BLE001 -- a registry that refuses: keep the alias only
    pass
