"""Training path of the image backbone + neck (the DHD-S / "DHD-B" configurations): mmdet
`ResNet(depth=50 | 101, style='pytorch', norm_eval=False, frozen_stages=-1)` (projects/configs/DHD/DHD-S.py:44-55 -- the
torchvision architecture; under the reference's runner it trains with BatchNorm on batch statistics, tools/train.py ->
mmdet3d train_model) and `CustomFPN` (projects/mmdet3d_plugin/models/necks/fpn.py:11-203), called from
BEVDet.image_encoder (detectors/bevdet.py:21-44).

Forward with saved activations and a hand-written backward on the same kernels as the rest of the training path
(dhd_b200/train.py): tcgen05 forward / data-gradient / weight-gradient GEMMs, batch-statistics BatchNorm from the
convolution epilogue + `dhd_bn_apply` (here with a bf16 identity path, `dhd_bn_apply_res16`), `dhd_act_bwd`,
`dhd_affine_combine`.  New pieces: the 7x7 stem as `dhd_stem_im2col` + a 1x1 GEMM whose weight gradient is folded back
to (64, 3, 7, 7), `dhd_maxpool3s2_bwd` (gather form, first-maximum rule, the stem's ReLU mask fused), the 1x1 stride-2
`downsample` convolutions (data gradient = one phase of the strided output grid), nearest up-sampling backward.

The residual stream of the Bottleneck stages stays bf16 (no fp32 side tensors), like the bf16 speed mode of
dhd_b200.backbone.  BatchNorm follows `train.BN_MODE` at construction ('batch' under module.train(), 'frozen' under
module.eval() with autograd on), exactly like the other trainers.
"""
import torch

from . import _lib
from . import dense as D
from .backbone import maxpool3s2, stem_im2col, upsample_nearest_add
from .modules import _p, _stream
from .train import _TrainConv, _acc, _ensure_grad, act_bwd


class _TrainStem(_TrainConv):
    """conv1 7x7 / 2 (+ bn1 + relu) as an im2col (K = (ky, kx, c), zero padded to a multiple of 64) and ONE 1x1 GEMM."""

    def __init__(self, conv, bn):
        w = conv.weight
        self.k, self.conv_stride, self.pad, self.cin_img = conv.kernel_size[0], conv.stride[0], conv.padding[0], w.shape[1]
        self.K = self.k * self.k * self.cin_img
        self.Kp = (self.K + 63) // 64 * 64
        self.weight, self.bias_p, self.bn = w, None, bn
        from . import train as T
        self.batch_bn = bn is not None and T.BN_MODE == 'batch'
        self.ksize, self.dilation, self.cols, self.stride = 1, 1, None, 1
        self.Cout, self.Cin, self.cin_pad, self.cout_pad = w.shape[0], self.Kp, self.Kp, w.shape[0]
        self.refresh()

    def refresh(self):
        from .modules import fold_bn
        w = self.weight
        if not hasattr(self, 'w_fwd'):
            if self.bn is not None and not self.batch_bn:
                s, b = fold_bn(self.bn, None)
                self.scale, self.bias = s.to(w.device), b.to(w.device)
            else:
                self.scale, self.bias = None, None
            self.w_fwd = torch.zeros(self.Cout, 1, 1, self.Kp, dtype=torch.bfloat16, device=w.device)
        self.w_fwd[:, 0, 0, :self.K].copy_(w.detach().permute(0, 2, 3, 1).reshape(self.Cout, self.K))

    def im2col(self, img):
        return stem_im2col(img.contiguous().float(), self.k, self.conv_stride, self.pad, 1)

    def backward(self, x, dy, dx_segs=None, bias_sums=None, **kw):
        """x: the im2col activation; the images need no gradient."""
        if self.batch_bn:
            dy = self._backward_batch_bn(dy)
        dw = D.conv2d_wgrad(x, dy, self.Cout, ksize=1, scale=self.scale)                 # (Cout, 1, Kp)
        g = dw[:, 0, :self.K].view(self.Cout, self.k, self.k, self.cin_img).permute(0, 3, 1, 2)
        _acc(self.weight, g)


class _TrainConvS2(_TrainConv):
    """1x1 stride-s `downsample` convolution of a stage's first Bottleneck.  Forward and weight gradient take the
    stride in the TMA box; the data gradient lands on the even pixels only (one phase), the rest is zero."""

    def __init__(self, weight, bn, stride):
        super().__init__(weight, None, bn, 1, stride=1)
        self.stride = stride

    def refresh(self):
        s, self.stride = self.stride, 1              # the parent's stride-2 branch builds 3x3 phase weights
        try:
            super().refresh()
        finally:
            self.stride = s

    def backward(self, x, dy, dx_segs=None, bias_sums=None, **kw):
        if self.stride == 1:
            return super().backward(x, dy, dx_segs, bias_sums, **kw)
        if self.batch_bn:
            dy = self._backward_batch_bn(dy)
        gw = _ensure_grad(self.weight)
        D.conv2d_wgrad(x, dy, self.Cout, ksize=1, scale=self.scale, stride=self.stride, grad=(gw, 0, self.Cin))
        if dx_segs is not None:
            out = dx_segs[0]['out_act']
            if x.H % 2 or x.W % 2 or out.coff != 0 or out.ld != out.C:
                raise NotImplementedError('stride-2 1x1 data gradient: even input size, a whole bf16 activation')
            out.data.zero_()
            ld = out.ld
            D.conv2d(dy, self.w_bwd, self.cin_pad, ksize=1, precision='bf16',
                     segs=[dict(out_act=out, out_view=(x.H * x.W * ld, 2 * x.W * ld, 2 * ld, 0))])


class _Buffers:
    def _act(self, name, N, H, W, C, zero=False):
        key = (name, N, H, W, C)
        if key not in self._buf:
            self._buf[key] = D.Act.empty(N, H, W, C, 1, self.device)
            if zero:
                self._buf[key].data.zero_()
        return self._buf[key]


class ImageResNetTrainer(_Buffers):
    """mmdet / torchvision ResNet-50 / -101 / -152 (Bottleneck, style='pytorch'): images -> the `out_indices` maps."""

    def __init__(self, net, device='cuda'):
        self.device = device
        self.stem = _TrainStem(net.conv1, net.bn1)
        self.layers = []
        for name in net.res_layers:
            blocks = []
            for b in getattr(net, name):
                s = b.conv2.stride[0]
                c1 = _TrainConv(b.conv1.weight, None, b.bn1, 1)
                c2 = _TrainConv(b.conv2.weight, None, b.bn2, 3, stride=s)
                c3 = _TrainConv(b.conv3.weight, None, b.bn3, 1)
                ds = None
                if b.downsample is not None:
                    ds = _TrainConvS2(b.downsample[0].weight, b.downsample[1], b.downsample[0].stride[0])
                blocks.append((c1, c2, c3, ds))
            self.layers.append(blocks)
        self.out_indices = tuple(net.out_indices)
        self.output_ids = list(self.out_indices)
        self._buf = {}

    def convs(self):
        yield self.stem
        for blocks in self.layers:
            for blk in blocks:
                for c in blk:
                    if c is not None:
                        yield c

    def refresh(self):
        for c in self.convs():
            c.refresh()

    def forward(self, img):
        """img: (N, 3, H, W) fp32 CUDA -> list of bf16 NHWC Acts (one per out_index)."""
        st = self.stem
        col = st.im2col(img)
        N = col.N
        s_out = self._act('stem', N, col.H, col.W, st.Cout)
        st.forward(col, [dict(act='relu', out_act=s_out)])
        x = maxpool3s2(s_out)
        self.saved_stem = (col, s_out)
        self.saved, outs = [], []
        for li, blocks in enumerate(self.layers):
            for bi, (c1, c2, c3, ds) in enumerate(blocks):
                tag = '%d_%d' % (li, bi)
                oH, oW = (x.H, x.W) if c2.stride == 1 else ((x.H + 1) // 2, (x.W + 1) // 2)
                t1 = self._act('t1_' + tag, N, x.H, x.W, c1.Cout)
                c1.forward(x, [dict(act='relu', out_act=t1)])
                t2 = self._act('t2_' + tag, N, oH, oW, c2.Cout)
                c2.forward(t1, [dict(act='relu', out_act=t2)])
                idn = x
                if ds is not None:
                    idn = self._act('id_' + tag, N, oH, oW, c3.Cout)
                    ds.forward(x, [dict(out_act=idn)])
                out = self._act('o_' + tag, N, oH, oW, c3.Cout)
                c3.forward(t2, [dict(act='relu', out_act=out)], residual_act=idn)
                self.saved.append((x, t1, t2, out))
                x = out
            if li in self.out_indices:
                outs.append(x)
        return outs

    def backward(self, dfeats):
        """dfeats: {layer index: Act gradient of that layer's output}; accumulates every parameter gradient (the images
        get none)."""
        idx, d = len(self.saved), None
        for li in range(len(self.layers) - 1, -1, -1):
            for bi in range(len(self.layers[li]) - 1, -1, -1):
                idx -= 1
                c1, c2, c3, ds = self.layers[li][bi]
                x, t1, t2, out = self.saved[idx]
                tag = '%d_%d' % (li, bi)
                extra = dfeats.get(li) if bi == len(self.layers[li]) - 1 else None
                if d is None:
                    if extra is None:
                        continue                     # layers behind the deepest used output carry no gradient
                    d = extra
                    act_bwd(d, out, 'relu')
                else:
                    act_bwd(d, out, 'relu', add=extra)
                # d = gradient at the block's pre-ReLU sum: goes to conv3's branch and to the identity / downsample
                dt2 = self._act('g_t2_' + tag, t2.N, t2.H, t2.W, t2.C)
                c3.backward(t2, d, [dict(out_act=dt2)])
                act_bwd(dt2, t2, 'relu')
                dt1 = self._act('g_t1_' + tag, t1.N, t1.H, t1.W, t1.C)
                c2.backward(t1, dt2, [dict(out_act=dt1)])
                act_bwd(dt1, t1, 'relu')
                # the identity / downsample path's gradient joins conv1's data gradient in that convolution's epilogue (one
                # rounding of the fp32 sum, no separate add pass)
                idg = d
                if ds is not None:
                    idg = self._act('g_id_' + tag, x.N, x.H, x.W, x.C)
                    ds.backward(x, d, [dict(out_act=idg)])
                dxin = self._act('g_x_' + tag, x.N, x.H, x.W, x.C)
                c1.backward(x, dt1, [dict(out_act=dxin)], residual_act=idg)
                d = dxin
        if d is None:
            return
        col, s_out = self.saved_stem
        ds_ = self._act('g_stem', s_out.N, s_out.H, s_out.W, s_out.C)
        _lib.check(_lib.load().dhd_maxpool3s2_bwd(_p(s_out.data), s_out.ld, s_out.coff, _p(d.data), d.ld, d.coff, s_out.N,
                                                  s_out.H, s_out.W, s_out.C, _p(ds_.data), ds_.ld, ds_.coff, 1, _stream()),
                   'maxpool3s2_bwd')
        self.stem.backward(col, ds_)


def _nearest_index(n_out, n_in, device):
    """F.interpolate(mode='nearest') source index per destination index (torch's rule, as dhd_upsample_nearest_add)."""
    i = torch.arange(n_out, device=device, dtype=torch.float32)
    return torch.clamp(torch.floor(i * (float(n_in) / float(n_out))).long(), max=n_in - 1)


class CustomFPNTrainer(_Buffers):
    """necks/fpn.py:153-203 for the DHD configurations (lateral 1x1 convs with bias, top-down nearest up-sampling + add,
    3x3 output conv on the `out_ids` levels; no norm, no activation, no extra levels): forward + backward."""

    def __init__(self, neck, device='cuda'):
        if neck.add_extra_convs or neck.num_outs > len(neck.out_ids):
            raise NotImplementedError('CustomFPN extra levels are not used by the DHD configs')
        self.device, self.start, self.out_ids = device, neck.start_level, list(neck.out_ids)
        self.out_channels = neck.out_channels
        self.lateral = [_TrainConv(m.conv.weight, m.conv.bias, None, 1) for m in neck.lateral_convs]
        self.fpn = [_TrainConv(m.conv.weight, m.conv.bias, None, 3) for m in neck.fpn_convs]
        self._buf = {}

    def refresh(self):
        for c in self.lateral + self.fpn:
            c.refresh()

    def forward(self, feats):
        """feats: list of Acts (backbone outputs, fine -> coarse) -> list of Acts (one per out_id)."""
        lats = []
        for i, conv in enumerate(self.lateral):
            f = feats[i + self.start]
            o = self._act('lat%d' % i, f.N, f.H, f.W, self.out_channels)
            conv.forward(f, [dict(out_act=o)])
            lats.append(o)
        for i in range(len(lats) - 1, 0, -1):
            upsample_nearest_add(lats[i], lats[i - 1])
        outs = []
        for j, i in enumerate(self.out_ids):
            o = self._act('out%d' % j, lats[i].N, lats[i].H, lats[i].W, self.out_channels)
            self.fpn[j].forward(lats[i], [dict(out_act=o)])
            outs.append(o)
        self.saved = ([feats[i + self.start] for i in range(len(self.lateral))], lats)
        return outs

    def backward(self, douts):
        """douts: list of Act gradients (one per out_id, None allowed) -> {input index: Act gradient}."""
        feats, lats = self.saved
        n = len(lats)
        dl = [None] * n                              # fp32 (N, H, W, C) gradients of the summed laterals
        for j, i in enumerate(self.out_ids):
            g = douts[j]
            if g is None:
                continue
            _, sums = act_bwd(g, None, None, want_sums=True)
            dx = self._act('g_lat%d' % i, lats[i].N, lats[i].H, lats[i].W, self.out_channels)
            self.fpn[j].backward(lats[i], g, [dict(out_act=dx)], bias_sums=sums[0])
            v = dx.data.float()
            dl[i] = v if dl[i] is None else dl[i] + v
        # lats[i-1] = own + nearest_up(lats[i]): the gradient flows down the pyramid, fine -> coarse
        for i in range(1, n):
            if dl[i - 1] is None:
                continue
            hi = dl[i - 1]
            N, H, W, C = hi.shape
            h, w = lats[i].H, lats[i].W
            if H == 2 * h and W == 2 * w:
                down = hi.view(N, h, 2, w, 2, C).sum(dim=(2, 4))
            else:
                tmp = torch.zeros(N, H, w, C, device=hi.device).index_add_(2, _nearest_index(W, w, hi.device), hi)
                down = torch.zeros(N, h, w, C, device=hi.device).index_add_(1, _nearest_index(H, h, hi.device), tmp)
            dl[i] = down if dl[i] is None else dl[i] + down
        grads = {}
        for i, conv in enumerate(self.lateral):
            if dl[i] is None:
                continue
            g = self._act('g_l%d' % i, lats[i].N, lats[i].H, lats[i].W, self.out_channels)
            g.data.copy_(dl[i])
            _, sums = act_bwd(g, None, None, want_sums=True)
            f = feats[i]
            dx = self._act('g_f%d' % i, f.N, f.H, f.W, f.C)
            conv.backward(f, g, [dict(out_act=dx)], bias_sums=sums[0])
            grads[i + self.start] = dx
        return grads
