"""Compiled (inference) forms of the BEV / voxel encoders between the voxel pool and the SFA fusion
(SURVEY.md 8(f) rank 1): the reference runs them as unfused cuDNN layers,
  UNet          models/backbones/unet.py:6-141   (DoubleConv, MaxPool2d(2), ConvTranspose2d(2, 2), pad, cat)
  CustomResNet  models/backbones/resnet.py:10-80 (mmdet BasicBlocks, first block of a stage stride 2 with a
                                                  3x3 stride-2 conv as `downsample`)
  FPN_LSS       models/necks/lss_fpn.py:11-74    (bilinear x4 up, cat, 2x conv3x3-BN-ReLU, bilinear x2 up,
                                                  conv3x3-BN-ReLU, conv1x1)
here every convolution is the tcgen05 implicit-GEMM kernel (stride-2 layers through a strided TMA box,
the four phases of the transposed convolution as 1x1 GEMMs that write a strided view), BatchNorm / bias /
ReLU / residual live in the epilogues, and concatenations are channel slices of one NHWC buffer that the
producers write in place -- no torch.cat, no F.pad, no layout copies."""
import ctypes

import os

import torch

from . import _lib
from . import dense as D
from .modules import _Conv, _p, _stream, fold_bn


def maxpool2(x, out):
    _lib.check(_lib.load().dhd_maxpool2(_p(x.data), x.ld, x.coff, x.part_stride, x.N, x.H, x.W, x.C, _p(out.data),
                                        out.ld, out.coff, out.part_stride, x.parts, _stream()), 'maxpool2')
    return out


def upsample_bilinear(x, out):
    """x -> out (an Act or slice whose N/H/W are the target size), align_corners=True."""
    _lib.check(_lib.load().dhd_upsample_bilinear(_p(x.data), x.ld, x.coff, x.part_stride, x.N, x.H, x.W, x.C, out.H,
                                                 out.W, _p(out.data), out.ld, out.coff, out.part_stride, x.parts,
                                                 _stream()), 'upsample_bilinear')
    return out


class _DoubleConv:
    """(conv3x3 -> BN -> ReLU) x 2, unet.py:45-61."""

    def __init__(self, seq, precision, device):
        self.c1 = _Conv(seq[0], seq[1], precision, device)
        self.c2 = _Conv(seq[3], seq[4], precision, device)
        self.mid, self.Cout = self.c1.Cout, self.c2.Cout

    def __call__(self, x, out, tmp):
        self.c1(x, [dict(act='relu', out_act=tmp)])
        self.c2(tmp, [dict(act='relu', out_act=out)])
        return out


class _ConvT2x2:
    """ConvTranspose2d(kernel 2, stride 2) (unet.py:86): out[2y+i, 2x+j] = W[:, :, i, j]^T x[y, x] + b, i.e. four
    1x1 GEMMs, each writing every second pixel of the output through a strided view."""

    def __init__(self, m, precision, device):
        parts = D.PRECISIONS[precision][0]
        w = m.weight.detach().float().to(device)                     # (Cin, Cout, 2, 2)
        self.Cout = w.shape[1]
        self.w = [D.pack_weight(w[:, :, i, j].t().contiguous(), parts) for i in range(2) for j in range(2)]
        self.bias = m.bias.detach().float().contiguous().to(device) if m.bias is not None else None
        self.precision = precision

    def __call__(self, x, out):
        """out: Act (or channel slice) of spatial size >= (2H, 2W); the up-sampled map lands at its top-left
        corner (the reference pads the remainder on the right / bottom: diff // 2 == 0 for diff <= 1)."""
        if out.H < 2 * x.H or out.W < 2 * x.W or out.H - 2 * x.H > 1 or out.W - 2 * x.W > 1:
            raise ValueError('ConvTranspose target %dx%d does not fit a %dx%d input' % (out.H, out.W, x.H, x.W))
        ld = out.ld
        for i in range(2):
            for j in range(2):
                D.conv2d(x, self.w[2 * i + j], self.Cout, precision=self.precision, bias=self.bias,
                         segs=[dict(out_act=out, out_view=(out.H * out.W * ld, 2 * out.W * ld, 2 * ld,
                                                           (i * out.W + j) * ld))])
        return out


class UNetEngine:
    def __init__(self, net, precision='fp32', device='cuda'):
        if net.bilinear:
            raise NotImplementedError('UNet(bilinear=True) is not used by the DHD configs')
        self.precision, self.parts, self.device = precision, D.PRECISIONS[precision][0], device
        self.inc = _DoubleConv(net.inc.double_conv, precision, device)
        self.down = [_DoubleConv(getattr(net, 'down%d' % k).maxpool_conv[1].double_conv, precision, device)
                     for k in range(1, 5)]
        self.upT = [_ConvT2x2(getattr(net, 'up%d' % k).up, precision, device) for k in range(1, 5)]
        self.upC = [_DoubleConv(getattr(net, 'up%d' % k).conv.double_conv, precision, device) for k in range(1, 5)]
        self.outc = _Conv(net.outc.conv, None, precision, device)
        self.n_classes = self.outc.Cout
        self._buf = {}

    def _act(self, name, N, H, W, C, zero=False):
        key = (name, N, H, W, C)
        if key not in self._buf:
            a = D.Act.empty(N, H, W, C, self.parts, self.device)
            if zero:
                a.data.zero_()
            self._buf[key] = a
        return self._buf[key]

    def __call__(self, x, out=None):
        """x: Act (B, n_channels, H, W) -> Act (B, n_classes, H, W) (written into `out` if given)."""
        N = x.N
        sizes = [(x.H, x.W)]
        for _ in range(4):
            sizes.append((sizes[-1][0] // 2, sizes[-1][1] // 2))
        chans = [self.inc.Cout] + [d.Cout for d in self.down]           # 64, 128, 256, 512, 1024
        # skip k and the up-sampled decoder map share one buffer: [skip | up]; the pad region (odd sizes) stays 0
        cats = [self._act('cat%d' % k, N, sizes[k][0], sizes[k][1], 2 * chans[k], zero=True) for k in range(4)]
        skip = [cats[k].slice(0, chans[k]) for k in range(4)]
        self.inc(x, skip[0], self._act('tmp0', N, sizes[0][0], sizes[0][1], self.inc.mid))
        cur = skip[0]
        for k in range(4):
            H, W = sizes[k + 1]
            pooled = self._act('pool%d' % k, N, H, W, chans[k])
            maxpool2(cur, pooled)
            dst = skip[k + 1] if k < 3 else self._act('bottom', N, H, W, chans[4])
            self.down[k](pooled, dst, self._act('tmpd%d' % k, N, H, W, self.down[k].mid))
            cur = dst
        for k in range(4):                                               # up1 .. up4
            lvl = 3 - k
            H, W = sizes[lvl]
            self.upT[k](cur, cats[lvl].slice(chans[lvl], 2 * chans[lvl]))
            dst = self._act('dec%d' % lvl, N, H, W, self.upC[k].Cout)
            self.upC[k](cats[lvl], dst, self._act('tmpu%d' % lvl, N, H, W, self.upC[k].mid))
            cur = dst
        if out is None:
            out = self._act('out', N, sizes[0][0], sizes[0][1], (self.n_classes + 63) // 64 * 64)
        self.outc(cur, [dict(out_act=out)])
        return out


class _ResBlock:
    """mmdet BasicBlock with an optional stride-2 first conv and conv `downsample` (resnet.py:47-55)."""

    def __init__(self, blk, precision, device):
        self.stride = blk.conv1.stride[0]
        self.c1 = _Conv(blk.conv1, blk.bn1, precision, device)
        self.c2 = _Conv(blk.conv2, blk.bn2, precision, device)
        self.ds = _Conv(blk.downsample, None, precision, device) if blk.downsample is not None else None
        self.Cout = self.c2.Cout

    def __call__(self, x, x32, new, new32):
        """x: Act, x32: its fp32 NHWC copy (identity path; None when the block has a downsample conv)."""
        oH, oW = (x.H, x.W) if self.stride == 1 else ((x.H + 1) // 2, (x.W + 1) // 2)
        nh = D.nhwc_strides(self.Cout, oH, oW)
        t = new(x.N, oH, oW, self.Cout)
        self.c1(x, [dict(act='relu', out_act=t)], stride=self.stride)
        if x.parts == 1 and os.environ.get('DHD_BF16_RESIDUAL', '1') != '0':
            # bf16 speed mode: the identity path is the bf16 activation itself / the downsample conv's bf16 output -- no
            # fp32 copy written by one block and re-read by the next (at DHD-L's 1024-channel 200x200 maps: 328 MB each)
            idn = x
            if self.ds is not None:
                idn = new(x.N, oH, oW, self.Cout)
                self.ds(x, [dict(out_act=idn)], stride=self.stride)
            out = new(x.N, oH, oW, self.Cout)
            self.c2(t, [dict(act='relu', out_act=out)], residual_act=idn)
            return out, None
        if self.ds is not None:
            idn = new32(x.N, oH, oW, self.Cout)
            self.ds(x, [dict(out_f32=(idn, nh))], stride=self.stride)
        else:
            idn = x32
        out, out32 = new(x.N, oH, oW, self.Cout), new32(x.N, oH, oW, self.Cout)
        self.c2(t, [dict(act='relu', out_act=out, out_f32=(out32, nh))], residual=(idn, nh[:3]))
        return out, out32


class CustomResNetEngine:
    def __init__(self, net, precision='fp32', device='cuda'):
        self.parts, self.device = D.PRECISIONS[precision][0], device
        self.stages = [[_ResBlock(b, precision, device) for b in stage] for stage in net.layers]
        self.output_ids = list(net.backbone_output_ids)

    def __call__(self, x):
        """x: Act (B, C, Dy, Dx) -> list of Acts (one per output stage)."""
        new = lambda N, H, W, C: D.Act.empty(N, H, W, C, self.parts, self.device)
        new32 = lambda N, H, W, C: torch.empty(N, H, W, C, device=self.device)
        feats, x32 = [], None
        for sid, stage in enumerate(self.stages):
            for blk in stage:
                x, x32 = blk(x, x32, new, new32)
            if sid in self.output_ids:
                feats.append(x)
        return feats


class FPNLSSEngine:
    def __init__(self, neck, precision='fp32', device='cuda'):
        if neck.lateral:
            raise NotImplementedError('FPN_LSS(lateral=...) is not used by the DHD configs')
        self.parts, self.device = D.PRECISIONS[precision][0], device
        self.idx = tuple(neck.input_feature_index)
        self.scale = int(neck.up.scale_factor)
        self.c1 = _Conv(neck.conv[0], neck.conv[1], precision, device)
        self.c2 = _Conv(neck.conv[3], neck.conv[4], precision, device)
        self.extra = neck.extra_upsample
        if self.extra:
            self.scale2 = int(neck.up2[0].scale_factor)
            self.c3 = _Conv(neck.up2[1], neck.up2[2], precision, device)
            self.c4 = _Conv(neck.up2[4], None, precision, device)
        self.out_channels = neck.out_channels

    def __call__(self, feats, out=None):
        """feats: list of Acts -> Act (B, out_channels, 2H, 2W) (written into `out` if given)."""
        new = lambda N, H, W, C: D.Act.empty(N, H, W, C, self.parts, self.device)
        x2, x1 = feats[self.idx[0]], feats[self.idx[1]]
        N, H, W = x2.N, x2.H, x2.W
        if (x1.H * self.scale, x1.W * self.scale) != (H, W):
            raise ValueError('FPN_LSS: %dx%d x%d does not match %dx%d' % (x1.H, x1.W, self.scale, H, W))
        cat = new(N, H, W, x2.C + x1.C)
        # [x2 | up(x1)]: the low-level map is copied into its slice by a 1:1 "up-sampling" (same size -> exact copy)
        upsample_bilinear(x2, cat.slice(0, x2.C))
        upsample_bilinear(x1, cat.slice(x2.C, x2.C + x1.C))
        t = new(N, H, W, self.c1.Cout)
        self.c1(cat, [dict(act='relu', out_act=t)])
        u = new(N, H, W, self.c2.Cout)
        self.c2(t, [dict(act='relu', out_act=u)])
        if not self.extra:
            return u
        H2, W2 = H * self.scale2, W * self.scale2
        v = new(N, H2, W2, u.C)
        upsample_bilinear(u, v)
        w = new(N, H2, W2, self.c3.Cout)
        self.c3(v, [dict(act='relu', out_act=w)])
        if out is None:
            out = new(N, H2, W2, self.c4.Cout)
        self.c4(w, [dict(out_act=out)])
        return out
