"""Compiled (inference) forms of the dense hot-path modules.

Each ``*Engine`` takes the torch module that owns the parameters (same parameter names as the
reference), folds eval-mode BatchNorm into per-channel scale / bias
vectors, packs the weights to split-bf16 once, and runs the forward pass entirely through the
C-ABI library: tcgen05 implicit-GEMM convolutions with fused epilogues plus a handful of
streaming kernels.  No torch.nn.functional call is on the forward path.

Reference forward passes restated here:
  HeightNet.forward     models/model_utils/depthnet.py:605-652 (trunk 430-484; ASPP 42-108;
                        Mlp 119-147; SELayer 150-169; mmdet BasicBlock; mmcv DeformConv2dPack)
  MGHS.depth_net        models/necks/lss_heightmap.py:62, 482-485
  SFA.forward           models/necks/mix.py:37-59, 87-90
  predictor.forward     models/dense_heads/occ_head.py:84-100
"""
import ctypes
import os

import torch

from . import _lib
from . import dense as D


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def fold_bn(bn, conv_bias=None):
    """Eval-mode BatchNorm2d/1d after a conv: y = conv*scale + bias."""
    scale = (bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps))
    bias = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    if conv_bias is not None:
        bias = bias + conv_bias.detach().float() * scale
    return scale.contiguous(), bias.contiguous()


def linear_rows(x, w, b=None, act=None, in_scale=None, in_shift=None, one_minus=False):
    """y = act(x' @ w.T + b) on fp32 rows through dhd_linear_rows (M = B*N rows: CUDA cores)."""
    R, K = x.shape
    O = w.shape[0]
    y = torch.empty(R, O, device=x.device)
    _lib.check(_lib.load().dhd_linear_rows(_p(x), R, K, _p(w), _p(b), O, D.ACT[act], _p(in_scale),
                                           _p(in_shift), int(one_minus), _p(y), _stream()), 'linear_rows')
    return y


def mean_hw(a):
    out = torch.empty(a.N, a.C, device=a.data.device)
    ws = D.mean_workspace(a.data.device, a.N, a.C, a.H * a.W)
    _lib.check(_lib.load().dhd_mean_hw(_p(a.data), a.ld, a.coff, a.part_stride, a.parts, a.N, a.C,
                                       a.H * a.W, _p(out), _p(ws), _stream()), 'mean_hw')
    return out


def unpack(a):
    """Act -> fp32 NCHW torch tensor (reference layout hand-off)."""
    out = torch.empty(a.N, a.C, a.H, a.W, device=a.data.device)
    _lib.check(_lib.load().dhd_unpack_nhwc_to_nchw(_p(a.data), a.ld, a.coff, a.part_stride, a.parts,
                                                   a.N, a.C, a.H, a.W, _p(out), _stream()), 'unpack')
    return out


class _Conv:
    """One packed convolution: weight parts + folded scale / bias."""

    def __init__(self, conv, bn, precision, device, weight=None, bias=None):
        parts = D.PRECISIONS[precision][0]
        w = conv.weight if weight is None else weight
        self.w = D.pack_weight(w.detach().to(device), parts)
        self.Cout = w.shape[0]
        self.ksize = w.shape[2] if w.dim() == 4 else 1
        self.dilation = conv.dilation[0] if hasattr(conv, 'dilation') and not isinstance(conv.dilation, int) \
            else getattr(conv, 'dilation', 1)
        cb = (conv.bias if bias is None else bias) if hasattr(conv, 'bias') else None
        if bn is not None:
            s, b = fold_bn(bn, cb)
            self.scale, self.bias = s.to(device), b.to(device)
        else:
            self.scale = None
            self.bias = cb.detach().float().contiguous().to(device) if cb is not None else None
        self.precision = precision

    def __call__(self, x, segs, **kw):
        return D.conv2d(x, self.w, self.Cout, ksize=self.ksize, dilation=self.dilation,
                        precision=self.precision, scale=self.scale, bias=self.bias, segs=segs, **kw)


class HeightNetEngine:
    """depthnet.py:418-487, 605-652, eval mode (HeightNet is always called with stereo_metas=None, lss_heightmap.py:787)."""

    def __init__(self, net, precision='fp32', device='cuda'):
        self.precision, self.device = precision, device
        self.parts = D.PRECISIONS[precision][0]
        dev = device
        f = lambda t: t.detach().float().contiguous().to(dev)
        self.C = net.reduce_conv[0].out_channels
        self.reduce = _Conv(net.reduce_conv[0], net.reduce_conv[1], precision, dev)
        # BatchNorm1d(27) folded into the first Linear's input transform
        self.bn_scale, self.bn_shift = [t.to(dev) for t in fold_bn(net.bn)]
        m = net.depth_mlp
        self.fc1_w, self.fc1_b, self.fc2_w, self.fc2_b = f(m.fc1.weight), f(m.fc1.bias), f(m.fc2.weight), f(m.fc2.bias)
        se = net.depth_se
        self.se_r_w, self.se_r_b = f(se.conv_reduce.weight.flatten(1)), f(se.conv_reduce.bias)
        self.se_e_w, self.se_e_b = f(se.conv_expand.weight.flatten(1)), f(se.conv_expand.bias)
        layers = list(net.depth_conv)
        self.blocks = []
        i = 0
        while i < len(layers) and type(layers[i]).__name__ == 'BasicBlock':
            blk = layers[i]
            w1, ds = blk.conv1.weight.detach().float(), None
            if blk.downsample is not None:
                # stereo DepthNet (depthnet.py:203-218): the first block reads cat(gated feature, cost volume) and
                # carries a plain 1x1 convolution on its identity path; both weights get zero columns up to the
                # 64-channel granule of the concatenation buffer
                cin = w1.shape[1]
                self.cat_channels = (cin + 63) // 64 * 64
                pad = self.cat_channels - cin
                wd = blk.downsample.weight.detach().float()
                w1 = torch.nn.functional.pad(w1, (0, 0, 0, 0, 0, pad))
                ds = _Conv(blk.downsample, None, precision, dev, weight=torch.nn.functional.pad(wd, (0, 0, 0, 0, 0, pad)))
            self.blocks.append((_Conv(blk.conv1, blk.bn1, precision, dev, weight=w1),
                                _Conv(blk.conv2, blk.bn2, precision, dev), ds))
            i += 1
        self.aspp = None
        if i < len(layers) and type(layers[i]).__name__ == 'ASPP':
            a = layers[i]
            mid = a.aspp1.atrous_conv.out_channels
            self.aspp_mid = mid
            self.aspp_branches = [_Conv(b.atrous_conv, b.bn, precision, dev) for b in (a.aspp1, a.aspp2, a.aspp3, a.aspp4)]
            # global-pool branch: mean -> 1x1 -> BN -> ReLU is a per-image vector; its share of
            # conv1 (columns [4*mid, 5*mid)) becomes a per-image bias of conv1's epilogue
            s5, b5 = fold_bn(a.global_avg_pool[2])
            self.gap_w = f(a.global_avg_pool[1].weight.flatten(1) * s5[:, None].to(a.global_avg_pool[1].weight.device))
            self.gap_b = b5.to(dev)
            s1, _ = fold_bn(a.bn1)
            w1 = a.conv1.weight.detach().float()
            self.aspp_out = _Conv(a.conv1, a.bn1, precision, dev, weight=w1[:, :4 * mid])
            self.aspp_w5 = f(w1[:, 4 * mid:].flatten(1) * s1[:, None].to(w1.device))
            self.aspp = a
            i += 1
        self.dcn = None
        if i < len(layers) and hasattr(layers[i], 'conv_offset'):
            dc = layers[i]
            self.dcn = dc
            self.dcn_groups = dc.groups
            self.dcn_offset = _Conv(dc.conv_offset, None, precision, dev)
            cg = self.C // dc.groups
            k = dc.weight.shape[2]
            self.dcn_k = k
            self.dcn_pad = dc.padding if isinstance(dc.padding, int) else dc.padding[0]
            self.dcn_dil = dc.dilation if isinstance(dc.dilation, int) else dc.dilation[0]
            # per group: [Cout/g][1 tap][parts][k*k*cg] with K ordered (tap, channel)
            self.dcn_w = []
            for g in range(dc.groups):
                wg = dc.weight.detach().float()[g * cg:(g + 1) * cg]                # (cg_out, cg, k, k)
                wg = wg.permute(0, 2, 3, 1).reshape(wg.shape[0], k * k * cg)        # K = (tap, c)
                self.dcn_w.append(D.pack_weight(wg.to(dev), self.parts))
            i += 1
        self.head = _Conv(layers[i], None, precision, dev)
        self.H_bins = self.head.Cout

    def gate(self, mlp_input):
        x = mlp_input.reshape(-1, mlp_input.shape[-1]).contiguous().float()
        h = linear_rows(x, self.fc1_w, self.fc1_b, 'relu', self.bn_scale, self.bn_shift)
        h = linear_rows(h, self.fc2_w, self.fc2_b)
        h = linear_rows(h, self.se_r_w, self.se_r_b, 'relu')
        return linear_rows(h, self.se_e_w, self.se_e_b, 'sigmoid')

    def __call__(self, x, mlp_input, softmax=True, hook=None, gate=None, batch_with=None):
        """x: Act (B*N, 256, fH, fW); returns height (B*N, H, fH, fW) fp32 NCHW (softmax-ed
        unless softmax=False, which gives the raw HeightNet output the reference returns).
        hook: optional callable run once after the ASPP branches are enqueued (the pipeline forks the
        geometry / binning kernels onto a side stream there, so they overlap the tail of this network).
        gate: the result of self.gate(mlp_input) when the caller computed it ahead (on another stream).
        batch_with: deferred convolutions (D.conv2d(..., defer=True)) that read other tensors and ride in the launch
        of the first BasicBlock convolution (the pipeline passes depth_net's 1x1 here)."""
        N, H, W, C, P, dev = x.N, x.H, x.W, self.C, self.parts, x.data.device
        new = lambda c: D.Act.empty(N, H, W, c, P, dev)
        nhwc = D.nhwc_strides(C, H, W)
        if gate is None:
            gate = self.gate(mlp_input)
        h = new(C)
        if P == 1 and os.environ.get('DHD_BF16_RESIDUAL', '1') != '0':
            self.reduce(x, [dict(act='relu', out_act=h)], img_gate=gate)
            return self.trunk(h, None, softmax, hook, batch_with)
        h32 = torch.empty(N, H, W, C, device=dev)
        self.reduce(x, [dict(act='relu', out_act=h, out_f32=(h32, nhwc))], img_gate=gate)
        return self.trunk(h, h32, softmax, hook, batch_with)

    def trunk(self, h, h32, softmax=True, hook=None, batch_with=None):
        """BasicBlocks -> ASPP -> DCN -> 1x1 head on the gated feature map (Act + its fp32 copy)."""
        N, H, W, C, P, dev = h.N, h.H, h.W, self.C, self.parts, h.data.device
        new = lambda c: D.Act.empty(N, H, W, c, P, dev)
        nhwc = D.nhwc_strides(C, H, W)
        lean = P == 1 and os.environ.get('DHD_BF16_RESIDUAL', '1') != '0'
        for c1, c2, ds in self.blocks:
            t = new(C)
            if lean:
                # bf16 speed mode: the identity path is the bf16 activation itself (no fp32 copy written by one block
                # and re-read by the next: 3 x 34 MB per block at DHD-S B=4) and every store is a 128-byte row
                idn = h
                if ds is not None:
                    idn = new(C)
                    ds(h, [dict(out_act=idn)])
                if batch_with:
                    D.conv2d_batch([c1(h, [dict(act='relu', out_act=t)], defer=True)] + list(batch_with))
                    batch_with = None
                else:
                    c1(h, [dict(act='relu', out_act=t)])
                h2 = new(C)
                c2(t, [dict(act='relu', out_act=h2)], residual_act=idn)
                h = h2
                continue
            if ds is not None:                       # identity path = 1x1 convolution of the concatenated input
                h32 = torch.empty(N, H, W, C, device=dev)
                ds(h, [dict(out_f32=(h32, nhwc))])
            if batch_with:
                D.conv2d_batch([c1(h, [dict(act='relu', out_act=t)], defer=True)] + list(batch_with))
                batch_with = None
            else:
                c1(h, [dict(act='relu', out_act=t)])
            h2, h2_32 = new(C), torch.empty(N, H, W, C, device=dev)
            c2(t, [dict(act='relu', out_act=h2, out_f32=(h2_32, nhwc))], residual=(h32, nhwc[:3]))
            h, h32 = h2, h2_32
        if batch_with:                               # no BasicBlock took them along
            D.conv2d_batch(list(batch_with))
        if self.aspp is not None:
            mid = self.aspp_mid
            cat = new(4 * mid)
            # the four branches read the same map and are 132 tiles each at DHD-S size: one persistent launch
            D.conv2d_batch([conv(h, [dict(act='relu', out_act=cat.slice(b * mid, (b + 1) * mid))], defer=True)
                            for b, conv in enumerate(self.aspp_branches)])
            if hook is not None:
                hook()
                hook = None
            x5 = linear_rows(mean_hw(h), self.gap_w, self.gap_b, 'relu')
            ib = linear_rows(x5, self.aspp_w5)
            h = new(C)
            self.aspp_out(cat, [dict(act='relu', out_act=h)], img_bias=ib)
        if self.dcn is not None:
            k, g = self.dcn_k, self.dcn_groups
            off = torch.empty(N, H, W, 2 * k * k, device=dev)
            self.dcn_offset(h, [dict(out_f32=(off, D.nhwc_strides(2 * k * k, H, W)))])
            col = new(k * k * C)
            _lib.check(_lib.load().dhd_dcn_im2col(
                _p(h.data), h.ld, h.coff, h.part_stride, h.parts, C, N, H, W, _p(off), 2 * k * k, k,
                self.dcn_pad, self.dcn_dil, g, _p(col.data), col.ld, col.part_stride, col.parts,
                _stream()), 'dcn_im2col')
            out = new(C)
            cg = C // g
            D.conv2d_batch([D.conv2d(col.slice(gi * k * k * cg, (gi + 1) * k * k * cg), self.dcn_w[gi], cg,
                                     precision=self.precision, defer=True,
                                     segs=[dict(out_act=out.slice(gi * cg, (gi + 1) * cg))]) for gi in range(g)])
            h = out
        if hook is not None:
            hook()
        height = torch.empty(N, self.H_bins, H, W, device=dev)
        self.head(h, [dict(act='softmax' if softmax else None,
                           out_f32=(height, D.nchw_strides(self.H_bins, H, W)))])
        return height


class DepthNetEngine(HeightNetEngine):
    """DepthNet of MGHS_Depth / MGHS_Stereo (depthnet.py:172-243, 362-415); with stereo=True the plane-sweep cost
    volume (dhd_b200.stereo) joins the depth trunk through cost_volumn_net.  The reduce_conv output feeds a camera-gated context branch (1x1 conv) and a camera-gated
    depth trunk (the same trunk as HeightNet).  Returns (depth (B*N, D, fH, fW) NCHW, softmax-ed
    or raw, context (B*N, fH, fW, C) NHWC) -- the pool's input layouts."""

    def __init__(self, net, precision='fp32', device='cuda'):
        super().__init__(net, precision, device)
        f = lambda t: t.detach().float().contiguous().to(device)
        m, se = net.context_mlp, net.context_se
        self.cfc1_w, self.cfc1_b, self.cfc2_w, self.cfc2_b = f(m.fc1.weight), f(m.fc1.bias), f(m.fc2.weight), f(m.fc2.bias)
        self.cse_r_w, self.cse_r_b = f(se.conv_reduce.weight.flatten(1)), f(se.conv_reduce.bias)
        self.cse_e_w, self.cse_e_b = f(se.conv_expand.weight.flatten(1)), f(se.conv_expand.bias)
        self.context = _Conv(net.context_conv, None, precision, device)
        self.Cctx = self.context.Cout
        self.stereo = bool(getattr(net, 'stereo', False))
        if self.stereo:
            # cost_volumn_net (depthnet.py:207-213): two stride-2 conv3x3 + BN on the D matching probabilities;
            # the D channels ride in a 64-channel granule (zero weight columns for the padding)
            cv = net.cost_volumn_net
            self.Dcv = cv[0].in_channels
            self.Dcv_pad = (self.Dcv + 63) // 64 * 64
            padw = lambda c: torch.nn.functional.pad(c.weight.detach().float(), (0, 0, 0, 0, 0, self.Dcv_pad - self.Dcv))
            self.cv1 = _Conv(cv[0], cv[1], precision, device, weight=padw(cv[0]))
            self.cv2 = _Conv(cv[2], cv[3], precision, device, weight=padw(cv[2]))
            self.cv_bias = float(net.bias)

    def context_gate(self, mlp_input):
        x = mlp_input.reshape(-1, mlp_input.shape[-1]).contiguous().float()
        h = linear_rows(x, self.cfc1_w, self.cfc1_b, 'relu', self.bn_scale, self.bn_shift)
        h = linear_rows(h, self.cfc2_w, self.cfc2_b)
        h = linear_rows(h, self.cse_r_w, self.cse_r_b, 'relu')
        return linear_rows(h, self.cse_e_w, self.cse_e_b, 'sigmoid')

    def _gated(self, x32, gate, want32):
        N, H, W, C = x32.shape
        out = D.Act.empty(N, H, W, C, self.parts, x32.device)
        o32 = torch.empty_like(x32) if want32 else None
        _lib.check(_lib.load().dhd_gate_channels(_p(x32), N, H * W, C, _p(gate), _p(out.data), out.ld, out.coff,
                                                 out.part_stride, out.parts, _p(o32), _stream()), 'gate_channels')
        return out, o32

    def new_cost_volume(self, N, H, W, device, zero=False):
        """The split-bf16 NHWC activation the cost-volume kernel fills (stereo map size H x W = 4fH x 4fW).
        zero=True: the all-zero volume the reference feeds when there is no previous frame (depthnet.py:389-396)."""
        a = D.Act.empty(N, H, W, self.Dcv_pad, self.parts, device)
        if zero:
            a.data.zero_()
        return a

    def __call__(self, x, mlp_input, softmax=True, cost_volume=None):
        N, H, W, C, dev = x.N, x.H, x.W, self.C, x.data.device
        x32 = torch.empty(N, H, W, C, device=dev)
        self.reduce(x, [dict(act='relu', out_f32=(x32, D.nhwc_strides(C, H, W)))])
        ctx, _ = self._gated(x32, self.context_gate(mlp_input), False)
        feat = torch.empty(N, H, W, self.Cctx, device=dev)
        self.context(ctx, [dict(out_f32=(feat, D.nhwc_strides(self.Cctx, H, W)))])
        if not self.stereo:
            lean = self.parts == 1 and os.environ.get('DHD_BF16_RESIDUAL', '1') != '0'     # trunk() then ignores h32
            h, h32 = self._gated(x32, self.gate(mlp_input), not lean)
            return self.trunk(h, h32, softmax), feat
        if cost_volume is None or (cost_volume.H + 3) // 4 != H or (cost_volume.W + 3) // 4 != W:
            raise ValueError('stereo DepthNet needs the (B*N, 4fH, 4fW, D) cost volume activation')
        # cat([gated depth feature, cost_volumn_net(cost volume)]) as two channel slices of one zero-padded buffer
        cat = D.Act(torch.zeros(N, H, W, self.parts * self.cat_channels, dtype=torch.bfloat16, device=dev),
                    self.cat_channels, self.parts)
        g = cat.slice(0, C)
        _lib.check(_lib.load().dhd_gate_channels(_p(x32), N, H * W, C, _p(self.gate(mlp_input)), _p(g.data), g.ld, g.coff,
                                                 g.part_stride, g.parts, None, _stream()), 'gate_channels')
        h2, w2 = (cost_volume.H + 1) // 2, (cost_volume.W + 1) // 2
        mid = D.Act(torch.zeros(N, h2, w2, self.parts * self.Dcv_pad, dtype=torch.bfloat16, device=dev),
                    self.Dcv_pad, self.parts)
        self.cv1(cost_volume, [dict(out_act=mid)], stride=2)
        self.cv2(mid, [dict(out_act=cat.slice(C, self.cat_channels))], stride=2)
        return self.trunk(cat, None, softmax), feat


class DepthHeadEngine:
    """MGHS.depth_net: one 1x1 conv -> softmax(depth) NCHW + context NHWC (the pool's layouts)."""

    def __init__(self, conv, n_depth, precision='fp32', device='cuda'):
        self.conv = _Conv(conv, None, precision, device)
        self.D = n_depth
        self.C = self.conv.Cout - n_depth

    def __call__(self, x, defer=False):
        """defer=True: returns (depth, feat, deferred convolution) -- the caller launches it inside a batch."""
        N, H, W, dev = x.N, x.H, x.W, x.data.device
        depth = torch.empty(N, self.D, H, W, device=dev)
        feat = torch.empty(N, H, W, self.C, device=dev)
        r = self.conv(x, [dict(c_lo=0, c_hi=self.D, act='softmax', out_f32=(depth, D.nchw_strides(self.D, H, W))),
                          dict(c_lo=self.D, c_hi=self.D + self.C, out_f32=(feat, D.nhwc_strides(self.C, H, W)))],
                      defer=defer)
        return (depth, feat, r) if defer else (depth, feat)


class SFAEngine:
    """mix.py:8-90 in eval mode."""

    def __init__(self, sfa, precision='fp32', device='cuda'):
        self.precision, self.parts = precision, D.PRECISIONS[precision][0]
        f = lambda t: t.detach().float().contiguous().to(device)
        st = sfa.mysk_7
        self.C = st.channels
        self.fc0_w, self.fc0_b = f(st.fc[0].weight), f(st.fc[0].bias)
        self.fc2_w, self.fc2_b = f(st.fc[2].weight), f(st.fc[2].bias)
        sl = st.spacial_leanring
        self.sp1 = _Conv(sl[0], sl[1], precision, device)
        self.sp1_w32 = f(sl[0].weight.flatten(1))           # [C][C] fp32: the gate is folded into it per call (_lean)
        self.sp2 = _Conv(sl[3], sl[4], precision, device)
        mr = sfa.mix_residual
        self.res1 = _Conv(mr[0], mr[1], precision, device)
        self.res2 = _Conv(mr[3], mr[4], precision, device)
        self.short = _Conv(sfa.mix_shortcut[0], sfa.mix_shortcut[1], precision, device)
        self.Cout = self.res2.Cout

    def _mix(self, x, a1, a2, out):
        _lib.check(_lib.load().dhd_sfa_mix(
            _p(x.data), x.ld, x.coff, x.part_stride, x.parts, self.C, x.N, x.H * x.W, _p(a1), _p(a2),
            _p(out.data), out.ld, out.coff, out.part_stride, out.parts, _stream()), 'sfa_mix')

    def __call__(self, x, out_f32=None):
        """x: Act (B, 2C, Dy, Dx) = cat(bev feature, voxel feature).  Returns Act (B, Cout, Dy, Dx)."""
        N, H, W, C, P, dev = x.N, x.H, x.W, self.C, self.parts, x.data.device
        new = lambda c: D.Act.empty(N, H, W, c, P, dev)
        s = x.mean if getattr(x, 'mean', None) is not None else mean_hw(x)
        a1 = linear_rows(linear_rows(s, self.fc0_w, self.fc0_b, 'relu'), self.fc2_w, self.fc2_b, 'sigmoid')
        if P == 1 and x.parts == 1 and os.environ.get('DHD_SFA_LEAN', '1') != '0':
            return self._lean(x, a1, out_f32)
        u = new(C)
        self._mix(x, a1, None, u)
        t = new(C)
        self.sp1(u, [dict(act='relu', out_act=t)])
        fuse = u                                   # reuse the buffer
        sc = torch.empty(N, H, W, self.Cout, device=dev)      # shortcut branch, fp32 NHWC
        if os.environ.get('DHD_SFA_FUSE', '0') != '0':      # measured 0.7 % slower than two passes: the 1x1 is epilogue-bound
            # spatial gate + second blend in ONE launch: the sigmoid gate never leaves the epilogue (conv desc mix_x)
            self.sp2(t, [dict(act='sigmoid', out_act=fuse)], sfa_mix=(x, a1))
        else:                                      # two-pass form (A/B measurements)
            self.sp2(t, [dict(act='sigmoid', out_f32=(sc, D.nhwc_strides(C, H, W)))])
            self._mix(x, a1, sc, fuse)
        self.short(x, [dict(out_f32=(sc, D.nhwc_strides(self.Cout, H, W)))])
        self.res1(fuse, [dict(act='relu', out_act=t)])
        out = new(self.Cout)
        seg = dict(act='relu', out_act=out)
        if out_f32 is not None:
            seg['out_f32'] = out_f32
        self.res2(t, [seg], residual=(sc, D.nhwc_strides(self.Cout, H, W)[:3]))
        return out


    def _lean(self, x, a1, out_f32):
        """bf16 speed mode: three memory passes less than the layer-by-layer form.
        (1) The channel-gated blend u = a1*bev + (1-a1)*vox is never materialised: it only feeds the 1x1 convolution
            spacial_leanring[0], and W u = [W diag(a1) | W diag(1-a1)] [bev ; vox], so the gate is folded into per-image
            weights (dhd_sfa_fold_gate, 4 x 256 x 512 values) and the convolution reads x directly.
        (2) The spatial gate a2 and the shortcut branch stay bf16 (128-byte-row TMA stores) instead of fp32 tensors.
        (3) The shortcut is added in mix_residual's last epilogue from its bf16 activation."""
        N, H, W, C, dev = x.N, x.H, x.W, self.C, x.data.device
        new = lambda c: D.Act.empty(N, H, W, c, 1, dev)
        lib = _lib.load()
        wq = torch.empty(N * C, 1, 1, 2 * C, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.dhd_sfa_fold_gate(_p(self.sp1_w32), _p(a1), N, C, C, _p(wq), _stream()), 'sfa_fold_gate')
        t = new(C)
        D.conv2d(x, wq, C, precision='bf16', scale=self.sp1.scale, bias=self.sp1.bias, image_weights=True,
                 segs=[dict(act='relu', out_act=t)])
        a2 = new(C)
        self.sp2(t, [dict(act='sigmoid', out_act=a2)])
        fuse = t                                   # t is dead once the gate exists
        _lib.check(lib.dhd_sfa_blend_b16(_p(x.data), x.ld, x.coff, C, N, H * W, _p(a1), _p(a2.data), a2.ld, a2.coff,
                                         _p(fuse.data), fuse.ld, fuse.coff, _stream()), 'sfa_blend_b16')
        sc = new(self.Cout)
        self.short(x, [dict(out_act=sc)])
        self.res1(fuse, [dict(act='relu', out_act=a2)])      # a2 is dead after the blend: reuse as res1's output
        out = t if self.Cout == C else new(self.Cout)
        seg = dict(act='relu', out_act=out)
        if out_f32 is not None:
            seg['out_f32'] = out_f32
        self.res2(a2, [seg], residual_act=sc)
        return out


class PredictorEngine:
    """occ_head.py:52-67, 84-100: 3x3 conv (+ mmcv ConvModule's default ReLU) -> permute ->
    Linear -> Softplus -> Linear -> (B, Dx, Dy, Dz, n_cls)."""

    def __init__(self, head, precision='fp32', device='cuda'):
        self.conv = _Conv(head.final_conv.conv, None, precision, device)
        self.relu = getattr(head.final_conv, 'with_activation', True)
        self.use_predicter = head.use_predicter
        if head.use_predicter:
            self.fc0 = _Conv(head.predicter[0], None, precision, device, weight=head.predicter[0].weight[:, :, None, None])
            self.fc2 = _Conv(head.predicter[2], None, precision, device, weight=head.predicter[2].weight[:, :, None, None])
        self.Dz, self.ncls = head.Dz, head.num_classes
        self.parts = D.PRECISIONS[precision][0]

    def fused_tail_ok(self):
        """The back-to-back GEMM kernel (dhd_predictor_tail) serves the bf16 speed mode of the DHD head shape
        (256 -> 512 -> 16 x 18); other shapes / the split-bf16 precision modes run layer by layer."""
        return (self.use_predicter and self.parts == 1 and self.conv.Cout == 256 and self.fc0.Cout == 512 and
                self.Dz == 16 and self.ncls == 18 and os.environ.get('DHD_TAIL_FUSED', '1') != '0')

    def __call__(self, x, occ=None, want_logits=True):
        """x: Act (B, C, Dy, Dx) -> occ_pred (B, Dx, Dy, Dz, n_cls) fp32 (None with want_logits=False).
        occ: optional uint8 (B, Dx, Dy, Dz) tensor that receives predictor.get_occ's class map
        (softmax(-1).argmax(-1), occ_head.py:141-153) -- in the fused kernel's epilogue when it runs, through
        dhd_occ_argmax otherwise."""
        N, H, W, P, dev = x.N, x.H, x.W, self.parts, x.data.device
        if not want_logits and occ is None:
            raise ValueError('nothing to compute: want_logits=False and no occ buffer')
        if not self.use_predicter:
            Co = self.conv.Cout
            out = torch.empty(N, W, H, Co, device=dev)
            self.conv(x, [dict(act='relu' if self.relu else None, out_f32=(out, (W * H * Co, Co, H * Co, 1)))])
            out = out.view(N, W, H, self.Dz, self.ncls)
            if occ is not None:
                occ_argmax(out, occ)
            return out
        t = D.Act.empty(N, H, W, self.conv.Cout, P, dev)
        self.conv(x, [dict(act='relu' if self.relu else None, out_act=t)])
        Co = self.fc2.Cout
        if self.fused_tail_ok():
            out = torch.empty(N, W, H, Co, device=dev) if want_logits else None
            D.predictor_tail(t, self.fc0.w, self.fc0.bias, self.fc2.w, self.fc2.bias, self.Dz, self.ncls,
                             logits=out, occ=occ)
            return out.view(N, W, H, self.Dz, self.ncls) if want_logits else None
        u = D.Act.empty(N, H, W, self.fc0.Cout, P, dev)
        self.fc0(t, [dict(act='softplus', out_act=u)])
        out = torch.empty(N, W, H, Co, device=dev)          # (B, Dx, Dy, C): permute(0, 3, 2, 1)
        self.fc2(u, [dict(out_f32=(out, (W * H * Co, Co, H * Co, 1)))])
        out = out.view(N, W, H, self.Dz, self.ncls)
        if occ is not None:
            occ_argmax(out, occ)
        return out


def occ_argmax(logits, occ):
    """occ[v] = argmax_k logits[v][k] (uint8), first maximum wins: predictor.get_occ, occ_head.py:141-153."""
    ncls = logits.shape[-1]
    if occ.dtype != torch.uint8 or not occ.is_contiguous() or not logits.is_contiguous() or \
            occ.numel() * ncls != logits.numel():
        raise ValueError('occ must be contiguous uint8 with one element per class row of the logits')
    _lib.check(_lib.load().dhd_occ_argmax(_p(logits), occ.numel(), ncls, _p(occ), _stream()), 'occ_argmax')
    return occ
