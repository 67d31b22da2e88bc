"""TEST INFRASTRUCTURE -- imports the *unmodified* reference hot-path modules from
/root/reference with sys.modules stubs for the un-vendored mmcv/mmdet/mmdet3d deps.

Only usable in the build container (``/root/reference`` does not exist on the GPU
box).  Used by ``oracle/make_golden.py`` and by the ``not gpu`` tests that pin the
oracle restatement (``oracle/mghs_oracle.py``) against the real reference code.
Nothing under ``dhd_b200/`` or ``projects/`` may import this file.

Stubbed symbols:
  mmcv.runner.BaseModule / force_fp32, mmcv.cnn.build_conv_layer / ConvModule,
  mmdet.models.backbones.resnet.BasicBlock, mmdet3d.models.builder.{NECKS,HEADS,
  build_loss}, and the relative import ``...ops.bev_pool_v2`` which is bound to the
  CPU restatement in ``oracle/mghs_oracle.py`` (the reference op is CUDA-only).
"""
import importlib.util
import os
import sys
import types
import warnings

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("DHD_REFERENCE_ROOT", "/root/reference")
_PLG = os.path.join(REF_ROOT, "projects", "mmdet3d_plugin")


def available():
    return os.path.isdir(_PLG)


class _Registry:
    def __init__(self):
        self.modules = {}

    def register_module(self, *a, **k):
        def deco(cls):
            self.modules[cls.__name__] = cls
            return cls
        return deco


class _BasicBlock(nn.Module):
    """Restatement of mmdet 2.25.1 ``BasicBlock`` (conv3x3-BN-ReLU-conv3x3-BN + id)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, **kw):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=dilation,
                               dilation=dilation, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return self.relu(out + identity)


class _ConvModule(nn.Module):
    """mmcv ConvModule: conv -> [norm] -> ReLU (mmcv's default act_cfg)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, bias=True, conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type='ReLU'), inplace=True, **kw):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding,
                              dilation, groups, bias=bool(bias) and norm_cfg is None)
        self.with_norm = norm_cfg is not None
        if self.with_norm:
            self.bn = nn.BatchNorm2d(out_channels)
        self.with_activation = act_cfg is not None
        if self.with_activation:
            self.activate = nn.ReLU(inplace=inplace)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def _build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    if typ == 'DCN':
        from .dense_oracle import DeformConv2dPack  # torchvision restatement
        cfg.pop('im2col_step', None)
        return DeformConv2dPack(*args, **cfg, **kwargs)
    if typ in ('Conv2d', None):
        return nn.Conv2d(*args, **cfg, **kwargs)
    raise KeyError(typ)


class _BaseModule(nn.Module):
    """mmcv.runner.BaseModule: an nn.Module whose constructor takes (and here ignores) init_cfg."""

    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


_LOADED = {}


def _install_stubs(bev_pool_v2_impl):
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    def force_fp32(*a, **k):
        def deco(fn):
            return fn
        return deco

    necks, heads, backbones = _Registry(), _Registry(), _Registry()
    mod('mmcv')
    mod('mmcv.runner', BaseModule=_BaseModule, force_fp32=force_fp32, auto_fp16=force_fp32)
    mod('mmcv.cnn', build_conv_layer=_build_conv_layer, ConvModule=_ConvModule,
        build_norm_layer=lambda cfg, n, postfix='': ('bn' + str(postfix), nn.BatchNorm2d(n)))
    mod('mmcv.cnn.bricks', ConvModule=_ConvModule).__path__ = []
    mod('mmcv.cnn.bricks.conv_module', ConvModule=_ConvModule)
    mod('mmdet')
    mod('mmdet.models', NECKS=necks)
    mod('mmdet.models.backbones')
    mod('mmdet.models.backbones.resnet', BasicBlock=_BasicBlock, Bottleneck=type('Bottleneck', (nn.Module,), {}))
    mod('mmdet3d')
    mod('mmdet3d.models', BACKBONES=backbones)
    mod('mmdet3d.models.builder', NECKS=necks, HEADS=heads, BACKBONES=backbones,
        build_loss=lambda cfg: None)
    # fake package tree so the reference's relative imports resolve
    for p in ('refplg', 'refplg.models', 'refplg.models.necks', 'refplg.models.model_utils',
              'refplg.models.dense_heads', 'refplg.models.losses', 'refplg.models.backbones'):
        m = mod(p)
        m.__path__ = []
    mod('refplg.ops', bev_pool_v2=bev_pool_v2_impl).__path__ = []
    return necks, heads


def _exec(modname, relpath):
    path = os.path.join(_PLG, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        spec.loader.exec_module(m)
    return m


def load_reference():
    """Returns a namespace with the reference classes: MGHS, MGHS_Depth, MGHS_Stereo,
    HeightNet, DepthNet, SFA, predictor (the real code from /root/reference)."""
    if 'ns' in _LOADED:
        return _LOADED['ns']
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    from . import mghs_oracle
    _install_stubs(mghs_oracle.bev_pool_v2)
    dn = _exec('refplg.models.model_utils.depthnet', 'models/model_utils/depthnet.py')
    mu = sys.modules['refplg.models.model_utils']
    mu.DepthNet, mu.HeightNet = dn.DepthNet, dn.HeightNet
    lh = _exec('refplg.models.necks.lss_heightmap', 'models/necks/lss_heightmap.py')
    mix = _exec('refplg.models.necks.mix', 'models/necks/mix.py')
    # occ_head imports ..losses.semkitti_loss (pure torch) relatively
    sk = _exec('refplg.models.losses.semkitti_loss', 'models/losses/semkitti_loss.py')
    oh = _exec('refplg.models.dense_heads.occ_head', 'models/dense_heads/occ_head.py')
    un = _exec('refplg.models.backbones.unet', 'models/backbones/unet.py')
    rn = _exec('refplg.models.backbones.resnet', 'models/backbones/resnet.py')
    fp = _exec('refplg.models.necks.lss_fpn', 'models/necks/lss_fpn.py')
    cf = _exec('refplg.models.necks.fpn', 'models/necks/fpn.py')
    ns = types.SimpleNamespace(
        UNet=un.UNet, CustomResNet=rn.CustomResNet, FPN_LSS=fp.FPN_LSS, CustomFPN=cf.CustomFPN,
        MGHS=lh.MGHS, MGHS_Depth=lh.MGHS_Depth, MGHS_Stereo=lh.MGHS_Stereo,
        HeightNet=dn.HeightNet, DepthNet=dn.DepthNet, ASPP=dn.ASPP, SFA=mix.SFA,
        predictor=oh.predictor, lss_heightmap=lh, depthnet=dn, mix=mix, occ_head=oh, semkitti_loss=sk)
    _LOADED['ns'] = ns
    return ns


def load_reference_configs():
    """Path of the reference's own config dir (tests load DHD-*.py from here)."""
    return os.path.join(REF_ROOT, 'projects', 'configs', 'DHD')
