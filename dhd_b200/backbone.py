"""Image backbone + neck of DHD-S / DHD-B on the tcgen05 convolution kernel (inference): the SURVEY 8(f)-4 widening.

Reference call sites: `img_backbone=dict(type='ResNet', depth=50 | 101, out_indices=(2, 3), style='pytorch')`
(projects/configs/DHD/DHD-S.py:44-55 -- mmdet 2.25.1 `ResNet`, the torchvision architecture with torchvision's
parameter names: `pretrained='torchvision://resnet50'`) and `img_neck=dict(type='CustomFPN', in_channels=[1024, 2048],
out_channels=256, num_outs=1, start_level=0, out_ids=[0])` (projects/mmdet3d_plugin/models/necks/fpn.py:11-203), called from
BEVDet.image_encoder (detectors/bevdet.py:21-44).

Layer map:
  conv1 7x7 / 2 (3 -> 64) + bn1 + relu   dhd_stem_im2col (K = 147 -> 192) + ONE 1x1 tcgen05 GEMM, BN / ReLU in its epilogue
  maxpool 3x3 / 2                         dhd_maxpool3s2
  Bottleneck (style='pytorch')            1x1 -> 3x3 (stride on the 3x3) -> 1x1 (+ 1x1 stride-s downsample), BN folded, the
                                          identity added in the last epilogue (bf16 speed mode: bf16 identity; split-bf16
                                          precision modes: fp32 identity), ReLU fused
  CustomFPN                               1x1 lateral convs (bias), += nearest up-sampling (dhd_upsample_nearest_add),
                                          3x3 output conv
This file is the inference form (eval-mode BatchNorm folded into the epilogues); the training form -- forward with saved
activations, batch-statistics BatchNorm and the hand-written backward -- is dhd_b200/train_backbone.py.
"""
import ctypes

import torch

from . import _lib
from . import dense as D
from .modules import _Conv, _p, _stream, fold_bn


def stem_im2col(img, ksize, stride, pad, parts):
    """(N, Cin, H, W) fp32 CUDA images -> Act (N, oH, oW, Kpad) with K = (ky, kx, c), zero padded to a multiple of 64."""
    N, Cin, H, W = img.shape
    K = ksize * ksize * Cin
    Kp = (K + 63) // 64 * 64
    oH, oW = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
    out = D.Act.empty(N, oH, oW, Kp, parts, img.device)          # the kernel writes the zero padding of every row too
    _lib.check(_lib.load().dhd_stem_im2col(_p(img), N, Cin, H, W, ksize, stride, pad, _p(out.data), out.ld, out.part_stride,
                                           parts, _stream()), 'stem_im2col')
    return out


def maxpool3s2(x):
    oH, oW = (x.H - 1) // 2 + 1, (x.W - 1) // 2 + 1
    out = D.Act.empty(x.N, oH, oW, x.C, x.parts, x.data.device)
    _lib.check(_lib.load().dhd_maxpool3s2(_p(x.data), x.ld, x.coff, x.part_stride, x.N, x.H, x.W, x.C, _p(out.data), out.ld,
                                          out.coff, out.part_stride, x.parts, _stream()), 'maxpool3s2')
    return out


def upsample_nearest_add(lo, io):
    """io += nearest-upsampled lo (in place)."""
    _lib.check(_lib.load().dhd_upsample_nearest_add(_p(lo.data), lo.ld, lo.coff, lo.part_stride, lo.H, lo.W, _p(io.data),
                                                    io.ld, io.coff, io.part_stride, io.N, io.H, io.W, io.C, io.parts,
                                                    _stream()), 'upsample_nearest_add')


class _Bottleneck:
    """mmdet / torchvision Bottleneck, style='pytorch' (the stride sits on the 3x3 conv2)."""

    def __init__(self, blk, precision, device):
        self.stride = blk.conv2.stride[0]
        self.c1 = _Conv(blk.conv1, blk.bn1, precision, device)
        self.c2 = _Conv(blk.conv2, blk.bn2, precision, device)
        self.c3 = _Conv(blk.conv3, blk.bn3, precision, device)
        self.ds = None
        if blk.downsample is not None:
            self.ds = _Conv(blk.downsample[0], blk.downsample[1], precision, device)
            self.ds_stride = blk.downsample[0].stride[0]
        self.mid, self.Cout = self.c1.Cout, self.c3.Cout
        self.parts = D.PRECISIONS[precision][0]

    def __call__(self, x, x32):
        """x: Act; x32: fp32 NHWC copy of x for the identity path of the split-bf16 modes (None in bf16 mode or when
        the block has a downsample branch).  Returns (Act, fp32 copy or None)."""
        N, dev, P = x.N, x.data.device, self.parts
        oH, oW = (x.H, x.W) if self.stride == 1 else ((x.H + 1) // 2, (x.W + 1) // 2)
        t1 = D.Act.empty(N, x.H, x.W, self.mid, P, dev)
        self.c1(x, [dict(act='relu', out_act=t1)])
        t2 = D.Act.empty(N, oH, oW, self.mid, P, dev)
        self.c2(t1, [dict(act='relu', out_act=t2)], stride=self.stride)
        out = D.Act.empty(N, oH, oW, self.Cout, P, dev)
        nh = D.nhwc_strides(self.Cout, oH, oW)
        if P == 1:                               # bf16 speed mode: bf16 identity, 128-byte-row stores, no fp32 side tensors
            idn = x
            if self.ds is not None:
                idn = D.Act.empty(N, oH, oW, self.Cout, 1, dev)
                self.ds(x, [dict(out_act=idn)], stride=self.ds_stride)
            self.c3(t2, [dict(act='relu', out_act=out)], residual_act=idn)
            return out, None
        if self.ds is not None:
            idn32 = torch.empty(N, oH, oW, self.Cout, device=dev)
            self.ds(x, [dict(out_f32=(idn32, nh))], stride=self.ds_stride)
        else:
            idn32 = x32
        out32 = torch.empty(N, oH, oW, self.Cout, device=dev)
        self.c3(t2, [dict(act='relu', out_act=out, out_f32=(out32, nh))], residual=(idn32, nh[:3]))
        return out, out32


class ResNetEngine:
    """mmdet `ResNet(depth=50 | 101 | 152, style='pytorch')` in eval mode: images -> the feature maps of `out_indices`."""

    def __init__(self, net, precision='bf16', device='cuda'):
        self.precision, self.parts, self.device = precision, D.PRECISIONS[precision][0], device
        c1 = net.conv1
        self.k, self.stride, self.pad = c1.kernel_size[0], c1.stride[0], c1.padding[0]
        cin = c1.in_channels
        K = self.k * self.k * cin
        self.Kp = (K + 63) // 64 * 64
        w = c1.weight.detach().float().permute(0, 2, 3, 1).reshape(c1.out_channels, K)          # K = (ky, kx, c)
        w = torch.nn.functional.pad(w, (0, self.Kp - K))[:, :, None, None]
        self.stem = _Conv(c1, net.bn1, precision, device, weight=w)
        self.stem.ksize, self.stem.dilation = 1, 1
        self.layers = [[_Bottleneck(b, precision, device) for b in getattr(net, name)] for name in net.res_layers]
        self.out_indices = tuple(net.out_indices)

    def __call__(self, img):
        """img: (N, 3, H, W) fp32 CUDA -> list of Acts (one per out_index)."""
        if not img.is_cuda:
            raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
        col = stem_im2col(img.contiguous().float(), self.k, self.stride, self.pad, self.parts)
        x = D.Act.empty(col.N, col.H, col.W, self.stem.Cout, self.parts, img.device)
        self.stem(col, [dict(act='relu', out_act=x)])
        x = maxpool3s2(x)
        x32 = None
        if self.parts > 1:                       # fp32 identity of the first block = the pooled stem output
            from .modules import unpack
            x32 = unpack(x).permute(0, 2, 3, 1).contiguous()
        outs = []
        for i, layer in enumerate(self.layers):
            for blk in layer:
                x, x32 = blk(x, x32)
            if i in self.out_indices:
                outs.append(x)
        return outs


class CustomFPNEngine:
    """necks/fpn.py:153-203 for the configurations DHD uses: lateral 1x1 convs on every input level, top-down nearest
    up-sampling + add, a 3x3 output conv on the levels in `out_ids` (no extra levels, no norm, no activation)."""

    def __init__(self, neck, precision='bf16', device='cuda'):
        if neck.add_extra_convs or neck.num_outs > len(neck.out_ids):
            raise NotImplementedError('CustomFPN extra levels are not used by the DHD configs')
        self.parts, self.device = D.PRECISIONS[precision][0], device
        self.start = neck.start_level
        self.lateral = [_Conv(m.conv, getattr(m, 'bn', None), precision, device) for m in neck.lateral_convs]
        self.fpn = [_Conv(m.conv, getattr(m, 'bn', None), precision, device) for m in neck.fpn_convs]
        self.relu = [getattr(m, 'with_activation', False) for m in list(neck.lateral_convs) + list(neck.fpn_convs)]
        if any(self.relu):
            raise NotImplementedError('CustomFPN(act_cfg=...) is not used by the DHD configs')
        self.out_ids = list(neck.out_ids)
        self.out_channels = neck.out_channels

    def __call__(self, feats):
        """feats: list of Acts (backbone outputs, fine -> coarse) -> list of Acts (one per out_id)."""
        new = lambda a, c: D.Act.empty(a.N, a.H, a.W, c, self.parts, self.device)
        lats = []
        for i, conv in enumerate(self.lateral):
            f = feats[i + self.start]
            o = new(f, self.out_channels)
            conv(f, [dict(out_act=o)])
            lats.append(o)
        for i in range(len(lats) - 1, 0, -1):
            upsample_nearest_add(lats[i], lats[i - 1])
        outs = []
        for j, i in enumerate(self.out_ids):
            o = new(lats[i], self.out_channels)
            self.fpn[j](lats[i], [dict(out_act=o)])
            outs.append(o)
        return outs
