"""The small slice of mmcv / mmdet / mmdet3d the DHD plugin API leans on, for environments
where those packages are absent (they are not installable offline): registries with
``register_module`` / ``build``, ``Config.fromfile`` with ``_base_`` inheritance, ``BaseModule``,
``force_fp32``, ``ConvModule``, ``build_conv_layer``, mmdet's ``BasicBlock`` and mmcv's
``DeformConv2dPack`` parameter containers.  When the real packages ARE importable their
registries are used instead, so the plugin registers into the genuine mmdet3d tables and
``projects/configs/DHD/DHD-*.py`` build through the stock ``build_model``.

Reference call sites: registry decorators lss_heightmap.py:12,704,900; mix.py:61; occ_head.py:32;
config loading tools/train.py:120-148; mmcv ConvModule occ_head.py:52-60; mmdet BasicBlock
depthnet.py:4,458-461; mmcv DCN depthnet.py:466-477.
"""
import copy
import os
import sys
import types

import torch
import torch.nn as nn


# ----------------------------------------------------------------------------- registries
class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force and self.module_dict[key] is not cls:
                raise KeyError('%s is already registered in %s' % (key, self.name))
            self.module_dict[key] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, **default_args):
        return build_from_cfg(cfg, self, default_args)

    def __contains__(self, key):
        return key in self.module_dict


def build_from_cfg(cfg, registry, default_args=None):
    if cfg is None:
        return None
    args = dict(cfg)
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    typ = args.pop('type')
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError('%s is not in the %s registry' % (typ, registry.name))
    return cls(**args)


try:   # genuine registries when mmdet3d is installed
    from mmdet3d.models.builder import (BACKBONES, DETECTORS, HEADS, LOSSES, NECKS,  # noqa: F401
                                        build_loss)
    HAVE_MMDET3D = True
except Exception:  # noqa: BLE001
    HAVE_MMDET3D = False
    BACKBONES, NECKS, HEADS = Registry('backbone'), Registry('neck'), Registry('head')
    DETECTORS, LOSSES = Registry('detector'), Registry('loss')

    def build_loss(cfg):
        return LOSSES.build(cfg) if cfg is not None and cfg.get('type') in LOSSES else None


def build_neck(cfg):
    return NECKS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_model(cfg, train_cfg=None, test_cfg=None):
    return DETECTORS.build(cfg, train_cfg=train_cfg, test_cfg=test_cfg)


# -------------------------------------------------------------------------------- modules
try:
    from mmcv.runner import BaseModule, force_fp32  # noqa: F401
except Exception:  # noqa: BLE001
    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

        def init_weights(self):
            pass

    def force_fp32(*a, **k):
        def deco(fn):
            return fn
        return deco


class EngineOwner:
    """Mixin of the plugin modules that run on a compiled engine (folded BatchNorm, packed bf16 weights).

    The engine is a SNAPSHOT of the parameters, so it is rebuilt whenever the module could have changed under it:
    a parameter / buffer modified in place (optimizer.step, `with no_grad(): p.copy_()` -- the tensor version
    counters), replaced or moved (`.to()`, `.cuda()`, `.half()`: `_apply`), reloaded (`load_state_dict`), or the
    module switched between train() and eval() (a repeated call with the same mode keeps them).  One blind spot remains: `p.data.copy_(...)` (mmcv's EMA hook) does
    not touch the version counter -- such writers are caught by the train()/eval() switch that follows them in the
    reference's runner, or call `invalidate_engines(model)`."""

    def _engine_state(self, device):
        import itertools
        return (str(device), bool(self.training)) + tuple(
            (id(t), t._version) for t in itertools.chain(self.parameters(), self.buffers()))

    def cached_engine(self, device, factory, slot='_engine'):
        key = self._engine_state(device)
        if self.__dict__.get(slot) is None or self.__dict__.get(slot + '__state') != key:
            self.__dict__[slot] = factory()
            self.__dict__[slot + '__state'] = key
        return self.__dict__[slot]

    def invalidate(self):
        for k in [k for k in self.__dict__ if k.endswith('_engine') or k == '_engine']:
            self.__dict__[k] = None

    def train(self, mode=True):
        # only a real switch drops the engines: a runner that calls model.eval() before every test batch must not pay
        # a rebuild (BatchNorm folding + weight packing of the whole model) per batch
        if bool(mode) != bool(self.training):
            self.invalidate()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self.invalidate()
        return super()._load_from_state_dict(*a, **k)


def invalidate_engines(model):
    """Drop every cached engine under `model` (after writing parameters through `.data`)."""
    for m in model.modules():
        if isinstance(m, EngineOwner):
            m.invalidate()


class ConvModule(nn.Module):
    """conv -> [norm] -> activation; mmcv's default act_cfg is ReLU (occ_head.py:52-60 relies on it).
    Parameter names: conv.{weight,bias}, bn.*"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'),
                 inplace=True, **kw):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                              bias=bool(bias))
        if self.with_norm:
            self.bn = nn.BatchNorm2d(out_channels)
        if self.with_activation:
            if act_cfg.get('type', 'ReLU') != 'ReLU':
                raise NotImplementedError('ConvModule activation %s' % act_cfg['type'])
            self.activate = nn.ReLU(inplace=inplace)


class BasicBlock(nn.Module):
    """Parameter container of mmdet 2.25.1 BasicBlock: conv1/bn1/conv2/bn2 (+ downsample)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, **kw):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=dilation, dilation=dilation,
                               bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample


class DeformConv2dPack(nn.Module):
    """Parameter container of mmcv DeformConv2dPack: weight (no bias) + zero-initialised
    conv_offset producing deform_groups*2*k*k channels."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deform_groups=1, bias=False, **kw):
        super().__init__()
        if stride != 1 or deform_groups != 1 or bias:
            raise NotImplementedError('DCN variant outside the DHD configs')
        k = kernel_size
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, k
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deform_groups = groups, deform_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, k, k))
        nn.init.kaiming_uniform_(self.weight, nonlinearity='relu')
        self.conv_offset = nn.Conv2d(in_channels, deform_groups * 2 * k * k, k, stride=stride,
                                     padding=padding, dilation=dilation, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)


def build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg or dict(type='Conv2d'))
    typ = cfg.pop('type')
    if typ == 'DCN':
        cfg.pop('im2col_step', None)
        return DeformConv2dPack(*args, **cfg, **kwargs)
    if typ in ('Conv2d', 'Conv', None):
        return nn.Conv2d(*args, **cfg, **kwargs)
    raise KeyError('conv layer type %s' % typ)


# --------------------------------------------------------------------------------- config
class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _to_cfgdict(o):
    if isinstance(o, dict):
        return ConfigDict((k, _to_cfgdict(v)) for k, v in o.items())
    if isinstance(o, (list, tuple)):
        return type(o)(_to_cfgdict(v) for v in o)
    return o


# what mmdet3d v1.0.0rc4's un-vendored _base_/default_runtime.py provides
_BASE_FALLBACK = dict(
    checkpoint_config=dict(interval=1),
    log_config=dict(interval=50, hooks=[dict(type='TextLoggerHook'), dict(type='TensorboardLoggerHook')]),
    dist_params=dict(backend='nccl'), log_level='INFO', work_dir=None, load_from=None,
    resume_from=None, workflow=[('train', 1)], opencv_num_threads=0, mp_start_method='fork')


def _merge(base, new):
    out = copy.deepcopy(base)
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get('_delete_', False):
            out[k] = _merge(out[k], v)
        else:
            if isinstance(v, dict):
                v = {a: b for a, b in v.items() if a != '_delete_'}
            out[k] = copy.deepcopy(v)
    return out


class Config:
    """Python-file configs with ``_base_`` inheritance and dotted overrides (mmcv.Config subset)."""

    def __init__(self, cfg_dict, filename=None):
        object.__setattr__(self, '_cfg_dict', _to_cfgdict(cfg_dict))
        object.__setattr__(self, 'filename', filename)

    @staticmethod
    def _file2dict(filename):
        filename = os.path.abspath(filename)
        src = open(filename).read()
        mod = types.ModuleType('_dhd_cfg_')
        mod.__file__ = filename
        exec(compile(src, filename, 'exec'), mod.__dict__)
        cfg = {k: v for k, v in mod.__dict__.items()
               if not k.startswith('__') and not isinstance(v, (types.ModuleType, types.FunctionType, type))}
        bases = cfg.pop('_base_', [])
        if isinstance(bases, str):
            bases = [bases]
        merged = {}
        for b in bases:
            path = os.path.join(os.path.dirname(filename), b)
            if os.path.exists(path):
                merged = _merge(merged, Config._file2dict(path))
            else:      # un-vendored mmdetection3d base: generic defaults the DHD configs override
                merged = _merge(merged, _BASE_FALLBACK)
        return _merge(merged, cfg)

    @staticmethod
    def fromfile(filename):
        return Config(Config._file2dict(filename), filename)

    def merge_from_dict(self, options):
        for key, v in options.items():
            d = self._cfg_dict
            ks = key.split('.')
            for k in ks[:-1]:
                d = d.setdefault(k, ConfigDict())
            d[ks[-1]] = _to_cfgdict(v)

    def __getattr__(self, k):
        return getattr(self._cfg_dict, k)

    def __getitem__(self, k):
        return self._cfg_dict[k]

    def __contains__(self, k):
        return k in self._cfg_dict

    def get(self, k, default=None):
        return self._cfg_dict.get(k, default)


def import_plugin(cfg, root=None):
    """tools/train.py:128-148: import the package named by cfg.plugin_dir so that its
    register_module decorators run."""
    import importlib
    if not cfg.get('plugin', False):
        return None
    d = os.path.dirname(cfg.plugin_dir.rstrip('/') + '/')
    mod = '.'.join(p for p in d.split('/') if p)
    if root is not None and root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module(mod)
