"""Training path of the dense hot-path layers: forward with saved activations, backward through
the C-ABI kernels (tcgen05 data / weight gradient GEMMs + the streaming kernels of csrc/train.cu),
gradients accumulated into the `.grad` of the torch parameters that own the weights -- so the
optimizer and the data-parallel gradient all-reduce (dhd_b200/shard.py) see ordinary parameters.

What the reference does here is torch autograd over cuDNN / cuBLAS (tools/train.py -> mmdet3d
train_model); the layers are the ones listed in dhd_b200/modules.py.  Arithmetic: bf16 operands,
fp32 accumulation (mixed-precision training), fp32 master weights and fp32 weight gradients.
BatchNorm layers run with frozen statistics and frozen affine parameters in this build
(FrozenBN fine-tuning): batch-statistics BatchNorm is the next step.
"""
import ctypes

import torch

from . import _lib
from . import dense as D
from .modules import _p, _stream, fold_bn

ACT_ID = {None: 0, 'none': 0, 'relu': 1, 'sigmoid': 2, 'softplus': 3}
_WS = {}


def _workspace(dev, nbytes):
    key = (dev, torch.cuda.current_stream().cuda_stream, 'act')
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WS[key] = ws
    return ws


def act_bwd(dy, y, act, out=None, want_sums=False, add=None):
    """dz = (dy [+ add]) * act'(y) (Acts, part 0).  out=None -> in place on dy.  Returns (dz Act, sums)
    with sums (2, C) fp32 = [sum_rows dz, sum_rows dz*y] when want_sums."""
    lib = _lib.load()
    C = dy.C
    out = dy if out is None else out
    sums = ws = None
    if want_sums:
        sums = torch.empty(2, C, device=dy.data.device)
        ws = _workspace(dy.data.device, lib.dhd_act_bwd_workspace_bytes(C))
    yy = y if y is not None else dy
    _lib.check(lib.dhd_act_bwd(_p(dy.data), dy.ld, dy.coff, _p(yy.data), yy.ld, yy.coff, dy.N * dy.H * dy.W, C,
                               ACT_ID[act], _p(out.data), out.ld, out.coff, _p(sums), _p(ws),
                               _p(add.data) if add is not None else None, add.ld if add is not None else 0,
                               add.coff if add is not None else 0, _stream()), 'act_bwd')
    return out, sums


def _acc(param, g):
    """param.grad += g (fp32), allocating on first use."""
    g = g.reshape(param.shape).to(param.dtype)
    if param.grad is None:
        param.grad = g.clone()
    else:
        param.grad.add_(g)


class _TrainConv:
    """One convolution of a trainable module: fp32 master weight lives in the torch parameter;
    `refresh()` re-packs the bf16 forward and data-gradient weights after an optimizer step."""

    def __init__(self, weight, bias, bn, ksize, dilation=1, cin_pad=None, cout_pad=None, cols=None):
        """cols=(lo, hi): the layer uses input-channel columns [lo, hi) of `weight` only (a branch of a
        concatenation whose other columns are handled elsewhere)."""
        self.weight, self.bias_p, self.bn = weight, bias, bn
        self.ksize, self.dilation, self.cols = ksize, dilation, cols
        self.Cout = weight.shape[0]
        self.Cin = weight.shape[1] if cols is None else cols[1] - cols[0]
        self.cin_pad = cin_pad or self.Cin          # forward input channels (zero padded to % 64)
        self.cout_pad = cout_pad or self.Cout       # backward input channels (dy padded to % 64)
        self.refresh()

    def refresh(self):
        """Re-pack the bf16 GEMM operands from the fp32 master weight: ONE dhd_pack_conv_weights launch.
        BatchNorm is frozen, so its folded scale is computed once; the folded bias is refreshed only where
        a trainable conv bias sits under the BatchNorm."""
        w = self.weight
        dev = w.device
        taps = self.ksize * self.ksize
        cin_total = w.shape[1]
        if not hasattr(self, 'w_fwd'):
            if self.bn is not None:
                s, b = fold_bn(self.bn, None)
                self.scale, self._bn_bias = s.to(dev), b.to(dev)
            else:
                self.scale = None
            self.w_fwd = torch.empty(self.Cout, taps, 1, self.cin_pad, dtype=torch.bfloat16, device=dev)
            self.w_bwd = torch.empty(self.Cin, taps, 1, self.cout_pad, dtype=torch.bfloat16, device=dev)
        if self.bn is None:
            self.bias = self.bias_p.detach() if self.bias_p is not None else None     # shares the parameter's storage
        else:
            self.bias = self._bn_bias if self.bias_p is None else self._bn_bias + self.bias_p.detach() * self.scale
        col_lo = 0 if self.cols is None else self.cols[0]
        _lib.check(_lib.load().dhd_pack_conv_weights(_p(w.detach()), self.Cout, cin_total, taps, col_lo, self.Cin,
                                                     _p(self.scale), _p(self.w_fwd), self.cin_pad, _p(self.w_bwd),
                                                     self.cout_pad, 0, _stream()), 'pack_conv_weights')

    def forward(self, x, segs, **kw):
        return D.conv2d(x, self.w_fwd, self.Cout, ksize=self.ksize, dilation=self.dilation, precision='bf16',
                        scale=self.scale, bias=self.bias, segs=segs, **kw)

    def backward(self, x, dy, dx_segs=None, bias_sums=None, **kw):
        """dy: Act = gradient w.r.t. this layer's pre-activation output (channels >= Cout zero).
        Accumulates weight / bias gradients; writes dx through `dx_segs` (conv2d segs) if given."""
        dw = D.conv2d_wgrad(x, dy, self.Cout, ksize=self.ksize, dilation=self.dilation, scale=self.scale)
        if x.C != self.Cin:
            dw = dw[:, :, :self.Cin]
        g = D.weight_grad_to_torch(dw.contiguous(), self.ksize)
        if self.cols is not None:
            _ensure_grad(self.weight)[:, self.cols[0]:self.cols[1]].add_(g)
        else:
            _acc(self.weight, g if self.weight.dim() == 4 else g[:, :, 0, 0])
        if self.bias_p is not None and bias_sums is not None:      # a conv bias under a frozen BN sees the BN scale
            _acc(self.bias_p, bias_sums[:self.Cout] if self.bn is None else bias_sums[:self.Cout] * self.scale)
        if dx_segs is not None:
            D.conv2d(dy, self.w_bwd, self.cin_pad, ksize=self.ksize, dilation=self.dilation, precision='bf16',
                     segs=dx_segs, **kw)


class PredictorTrainer:
    """predictor (occ_head.py:52-67, 84-131): forward, class-weighted masked cross-entropy, backward."""

    def __init__(self, head, device='cuda'):
        if not head.use_predicter:
            raise NotImplementedError('use_predicter=False head')
        self.head, self.device = head, device
        self.Dz, self.ncls = head.Dz, head.num_classes
        self.conv = _TrainConv(head.final_conv.conv.weight, head.final_conv.conv.bias, None, 3)
        self.relu = getattr(head.final_conv, 'with_activation', True)
        nout = self.Dz * self.ncls
        self.nout, self.nout_pad = nout, (nout + 63) // 64 * 64
        self.fc0 = _TrainConv(head.predicter[0].weight, head.predicter[0].bias, None, 1)
        self.fc2 = _TrainConv(head.predicter[2].weight, head.predicter[2].bias, None, 1, cout_pad=self.nout_pad)
        cw = getattr(head, 'cls_weights', None) if head.class_balance else None
        self.class_weight = cw.float().to(device).contiguous() if cw is not None else None
        self.loss_weight = float(getattr(head.loss_occ, 'loss_weight', 1.0)) * float(head.weight_ce) \
            if head.loss_occ is not None else float(head.weight_ce)
        self.ignore_index = int(getattr(head.loss_occ, 'ignore_index', 255) or 255) if head.loss_occ is not None else 255
        # sem_scal / geo_scal terms of predictor.loss (occ_head.py:129-131); free class = num_classes - 1
        self.weight_sem, self.weight_geo = float(head.weight_sem), float(head.weight_geo)
        self._buf = {}

    def refresh(self):
        for c in (self.conv, self.fc0, self.fc2):
            c.refresh()

    def _act(self, name, N, H, W, C, zero=False):
        key = (name, N, H, W, C)
        a = self._buf.get(key)
        if a is None:
            a = D.Act.empty(N, H, W, C, 1, self.device)
            if zero:
                a.data.zero_()
            self._buf[key] = a
        return a

    def forward(self, x):
        """x: Act (B, C, Dy, Dx) -> occ_pred (B, Dx, Dy, Dz, n_cls) fp32; activations saved."""
        N, H, W = x.N, x.H, x.W
        t = self._act('t', N, H, W, self.conv.Cout)
        self.conv.forward(x, [dict(act='relu' if self.relu else None, out_act=t)])
        u = self._act('u', N, H, W, self.fc0.Cout)
        self.fc0.forward(t, [dict(act='softplus', out_act=u)])
        Co = self.nout
        out = torch.empty(N, W, H, Co, device=self.device)
        self.fc2.forward(u, [dict(out_f32=(out, (W * H * Co, Co, H * Co, 1)))])
        self.saved = (x, t, u, out)
        return out.view(N, W, H, self.Dz, self.ncls)

    def loss(self, voxel_semantics, mask_camera=None, scal=True):
        """predictor.loss on the logits of the last forward: returns a 4-element device tensor
        [loss_occ, avg_factor, loss_voxel_sem_scal, loss_voxel_geo_scal] (the scal terms are 0 with scal=False);
        the gradient of their sum w.r.t. the logits is kept for backward()."""
        x, t, u, out = self.saved
        N, H, W = x.N, x.H, x.W
        labels = voxel_semantics.to(torch.uint8).contiguous()
        mask = mask_camera.to(torch.uint8).contiguous() if mask_camera is not None else None
        dlog = self._act('dlog', N, H, W, self.nout_pad, zero=True)
        res = torch.empty(4, device=self.device)
        lib = _lib.load()
        ws = _workspace(self.device, lib.dhd_occ_loss_workspace_bytes())
        _lib.check(lib.dhd_occ_ce_loss(_p(out), _p(labels), _p(mask), _p(self.class_weight), self.ncls,
                                       self.ignore_index, N, W, H, self.Dz, self.loss_weight,
                                       self.weight_sem if scal else 0.0, self.weight_geo if scal else 0.0, self.ncls - 1,
                                       _p(res), _p(dlog.data), dlog.ld, _p(ws), _stream()), 'occ_ce_loss')
        self.dlog = dlog
        return res

    def backward(self, want_dx=True):
        """Backward of loss() through the head.  Returns dL/dx as an Act (bf16) or None."""
        x, t, u, out = self.saved
        N, H, W = x.N, x.H, x.W
        dlog = self.dlog
        _, sums = act_bwd(dlog, None, None, want_sums=True)                   # bias gradient of predicter[2]
        du = self._act('du', N, H, W, self.fc0.Cout)
        self.fc2.backward(u, dlog, [dict(out_act=du)], bias_sums=sums[0])
        _, sums = act_bwd(du, u, 'softplus', want_sums=True)
        dt = self._act('dt', N, H, W, self.conv.Cout)
        self.fc0.backward(t, du, [dict(out_act=dt)], bias_sums=sums[0])
        _, sums = act_bwd(dt, t, 'relu' if self.relu else None, want_sums=True)
        dx = self._act('dx', N, H, W, x.C) if want_dx else None
        self.conv.backward(x, dt, [dict(out_act=dx)] if want_dx else None, bias_sums=sums[0])
        return dx


class DepthHeadTrainer:
    """MGHS.depth_net (lss_heightmap.py:62, 482-489): 1x1 conv -> softmax(depth) | context, and its
    backward from the pool's depth / context gradients."""

    def __init__(self, conv, n_depth, device='cuda'):
        self.D = n_depth
        self.C = conv.weight.shape[0] - n_depth
        self.device = device
        self.cout_pad = (conv.weight.shape[0] + 63) // 64 * 64
        self.conv = _TrainConv(conv.weight, conv.bias, None, 1, cout_pad=self.cout_pad)
        self._dy = None

    def refresh(self):
        self.conv.refresh()

    def forward(self, x):
        N, H, W = x.N, x.H, x.W
        depth = torch.empty(N, self.D, H, W, device=self.device)
        feat = torch.empty(N, H, W, self.C, device=self.device)
        self.conv.forward(x, [dict(c_lo=0, c_hi=self.D, act='softmax', out_f32=(depth, D.nchw_strides(self.D, H, W))),
                              dict(c_lo=self.D, c_hi=self.D + self.C, out_f32=(feat, D.nhwc_strides(self.C, H, W)))])
        self.saved = (x, depth)
        return depth, feat

    def backward(self, depth_grad, feat_grad, want_dx=True):
        """depth_grad (BN, D, fH, fW), feat_grad (.., fH, fW, C) fp32 from dhd_mghs_pool_bwd."""
        x, depth = self.saved
        N, H, W = x.N, x.H, x.W
        if self._dy is None or (self._dy.N, self._dy.H, self._dy.W) != (N, H, W):
            self._dy = D.Act.empty(N, H, W, self.cout_pad, 1, self.device)
        dy = self._dy
        _lib.check(_lib.load().dhd_depth_head_bwd(_p(depth), _p(depth_grad.contiguous()), _p(feat_grad.contiguous()),
                                                  N, self.D, H * W, self.C, _p(dy.data), dy.ld, _stream()),
                   'depth_head_bwd')
        _, sums = act_bwd(dy, None, None, want_sums=True)
        dx = D.Act.empty(N, H, W, x.C, 1, self.device) if want_dx else None
        self.conv.backward(x, dy, [dict(out_act=dx)] if want_dx else None, bias_sums=sums[0])
        return dx


class SFATrainer:
    """SFA (mix.py:8-90) with frozen BatchNorm: forward with saved activations and the backward of
    every convolution, both gated blends and the squeeze MLP."""

    def __init__(self, sfa, device='cuda'):
        from .modules import linear_rows, mean_hw
        self._linear, self._mean = linear_rows, mean_hw
        self.sfa, self.device = sfa, device
        st = sfa.mysk_7
        self.C = st.channels
        self.fc0, self.fc2 = st.fc[0], st.fc[2]
        sl, mr, ms = st.spacial_leanring, sfa.mix_residual, sfa.mix_shortcut
        self.sp1 = _TrainConv(sl[0].weight, sl[0].bias, sl[1], 1)
        self.sp2 = _TrainConv(sl[3].weight, sl[3].bias, sl[4], 1)
        self.res1 = _TrainConv(mr[0].weight, None, mr[1], 3)
        self.res2 = _TrainConv(mr[3].weight, None, mr[4], 3)
        self.short = _TrainConv(ms[0].weight, None, ms[1], 1)
        self.Cout = self.res2.Cout
        self._buf = {}

    def refresh(self):
        for c in (self.sp1, self.sp2, self.res1, self.res2, self.short):
            c.refresh()

    def _act(self, name, N, H, W, C):
        key = (name, N, H, W, C)
        if key not in self._buf:
            self._buf[key] = D.Act.empty(N, H, W, C, 1, self.device)
        return self._buf[key]

    def _f32(self, name, *shape):
        key = (name,) + shape
        if key not in self._buf:
            self._buf[key] = torch.empty(*shape, device=self.device)
        return self._buf[key]

    def _mix(self, x, a1, a2, out):
        _lib.check(_lib.load().dhd_sfa_mix(_p(x.data), x.ld, x.coff, x.part_stride, x.parts, self.C, x.N, x.H * x.W,
                                           _p(a1), _p(a2), _p(out.data), out.ld, out.coff, out.part_stride, out.parts,
                                           _stream()), 'sfa_mix')

    def forward(self, x):
        """x: Act (B, 2C, Dy, Dx) = cat(bev feature, voxel feature) -> Act (B, Cout, Dy, Dx)."""
        N, H, W, C = x.N, x.H, x.W, self.C
        f = lambda t: t.detach().float().contiguous()
        s = x.mean if getattr(x, 'mean', None) is not None else self._mean(x)
        h = self._linear(s, f(self.fc0.weight), f(self.fc0.bias), 'relu')
        a1 = self._linear(h, f(self.fc2.weight), f(self.fc2.bias), 'sigmoid')
        u = self._act('u', N, H, W, C)
        self._mix(x, a1, None, u)
        t = self._act('t', N, H, W, C)
        self.sp1.forward(u, [dict(act='relu', out_act=t)])
        a2 = self._f32('a2', N, H, W, C)
        self.sp2.forward(t, [dict(act='sigmoid', out_f32=(a2, D.nhwc_strides(C, H, W)))])
        fuse = self._act('fuse', N, H, W, C)
        self._mix(x, a1, a2, fuse)
        sc = self._f32('sc', N, H, W, self.Cout)
        self.short.forward(x, [dict(out_f32=(sc, D.nhwc_strides(self.Cout, H, W)))])
        r = self._act('r', N, H, W, self.Cout)
        self.res1.forward(fuse, [dict(act='relu', out_act=r)])
        out = self._act('out', N, H, W, self.Cout)
        self.res2.forward(r, [dict(act='relu', out_act=out)], residual=(sc, D.nhwc_strides(self.Cout, H, W)[:3]))
        self.saved = (x, s, h, a1, u, t, a2, fuse, r, out)
        return out

    def backward(self, dout, want_dx=True):
        """dout: Act = dL/d(SFA output) (modified in place).  Returns dL/dx as an Act (bf16)."""
        x, s, h, a1, u, t, a2, fuse, r, out = self.saved
        N, H, W, C = x.N, x.H, x.W, self.C
        lib = _lib.load()
        act_bwd(dout, out, 'relu')
        dr = self._act('dr', N, H, W, self.Cout)
        self.res2.backward(r, dout, [dict(out_act=dr)])
        act_bwd(dr, r, 'relu')
        dfuse = self._act('dfuse', N, H, W, C)
        self.res1.backward(fuse, dr, [dict(out_act=dfuse)])
        dxa = self._f32('dxa', N, H, W, 2 * C)
        dxb = self._f32('dxb', N, H, W, 2 * C)
        dpre2 = self._act('dpre2', N, H, W, C)
        S = self._f32('S', N, C)
        ws = _workspace(self.device, lib.dhd_sfa_gate_bwd_workspace_bytes(N, H * W, C))
        _lib.check(lib.dhd_sfa_gate_bwd(0, _p(dfuse.data), dfuse.ld, dfuse.coff, _p(x.data), x.ld, x.coff, C, N, H * W,
                                        _p(a1), _p(a2), _p(dpre2.data), dpre2.ld, dpre2.coff, _p(dxa), _p(S), 0,
                                        _p(ws), _stream()), 'sfa_gate_bwd(fuse)')
        # shortcut: weight gradient, and dxb = dxa + dgrad(dout) through the residual input of the epilogue
        st2 = D.nhwc_strides(2 * C, H, W)
        self.short.backward(x, dout, [dict(out_f32=(dxb, st2))], residual=(dxa, st2[:3]))
        dt = self._act('dt', N, H, W, C)
        _, sums = act_bwd(dpre2, None, None, want_sums=True)
        self.sp2.backward(t, dpre2, [dict(out_act=dt)], bias_sums=sums[0])
        _, sums = act_bwd(dt, t, 'relu', want_sums=True)
        du = self._act('du', N, H, W, C)
        self.sp1.backward(u, dt, [dict(out_act=du)], bias_sums=sums[0])
        _lib.check(lib.dhd_sfa_gate_bwd(1, _p(du.data), du.ld, du.coff, _p(x.data), x.ld, x.coff, C, N, H * W,
                                        _p(a1), None, None, 0, 0, _p(dxb), _p(S), 1, _p(ws), _stream()),
                   'sfa_gate_bwd(u)')
        # squeeze MLP (B rows): CUDA-core products through dhd_linear_rows
        lin = self._linear
        dz2 = (S * a1 * (1.0 - a1)).contiguous()
        _acc(self.fc2.weight, lin(dz2.t().contiguous(), h.t().contiguous()))
        _acc(self.fc2.bias, dz2.sum(0))
        dh = lin(dz2, self.fc2.weight.detach().float().t().contiguous())
        dz0 = (dh * (h > 0).float()).contiguous()
        _acc(self.fc0.weight, lin(dz0.t().contiguous(), s.t().contiguous()))
        _acc(self.fc0.bias, dz0.sum(0))
        if not want_dx:
            return None
        ds = lin(dz0, self.fc0.weight.detach().float().t().contiguous()) * (1.0 / (H * W))
        dx = self._act('dx', N, H, W, 2 * C)
        _lib.check(lib.dhd_add_rowvec(_p(dxb), _p(ds.contiguous()), N, H * W, 2 * C, _p(dx.data), dx.ld, dx.coff,
                                      _stream()), 'add_rowvec')
        return dx


def _ensure_grad(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


class HeightNetTrainer:
    """HeightNet (depthnet.py:418-487, 605-652, non-stereo) with frozen BatchNorm and Dropout off:
    reduce conv + camera-aware SE gate, BasicBlocks, ASPP (global branch as a per-image bias), DCN
    (deformable im2col + grouped GEMM), 1x1 head + softmax; backward of all of it, fed by the height loss
    (lss_heightmap.py:595-622)."""

    def __init__(self, net, device='cuda', loss_weight=0.1):
        from .modules import linear_rows, mean_hw
        self._linear, self._mean = linear_rows, mean_hw
        self.net, self.device, self.loss_weight = net, device, float(loss_weight)
        self.C = net.reduce_conv[0].out_channels
        self.reduce = _TrainConv(net.reduce_conv[0].weight, net.reduce_conv[0].bias, net.reduce_conv[1], 3)
        layers = list(net.depth_conv)
        self.blocks, i = [], 0
        while i < len(layers) and type(layers[i]).__name__ == 'BasicBlock':
            b = layers[i]
            if b.downsample is not None:
                raise NotImplementedError('stereo downsample branch')
            self.blocks.append((_TrainConv(b.conv1.weight, None, b.bn1, 3), _TrainConv(b.conv2.weight, None, b.bn2, 3)))
            i += 1
        a = layers[i]
        if type(a).__name__ != 'ASPP':
            raise NotImplementedError('HeightNet without ASPP')
        self.aspp = a
        self.mid = mid = a.aspp1.atrous_conv.out_channels
        self.branches = [_TrainConv(b.atrous_conv.weight, None, b.bn, b.atrous_conv.kernel_size[0], b.atrous_conv.dilation[0])
                         for b in (a.aspp1, a.aspp2, a.aspp3, a.aspp4)]
        self.aspp_out = _TrainConv(a.conv1.weight, None, a.bn1, 1, cols=(0, 4 * mid))
        i += 1
        self.dcn = layers[i] if hasattr(layers[i], 'conv_offset') else None
        if self.dcn is None:
            raise NotImplementedError('HeightNet without DCN (DHD-L): use_dcn=False trunk')
        dc = self.dcn
        self.k, self.groups = dc.weight.shape[2], dc.groups
        self.pad = dc.padding if isinstance(dc.padding, int) else dc.padding[0]
        self.dil = dc.dilation if isinstance(dc.dilation, int) else dc.dilation[0]
        self.noff = 2 * self.k * self.k
        self.dcn_offset = _TrainConv(dc.conv_offset.weight, dc.conv_offset.bias, None, self.k, cout_pad=64)
        i += 1
        head = layers[i]
        self.H_bins = head.weight.shape[0]
        self.head = _TrainConv(head.weight, head.bias, None, 1, cout_pad=(self.H_bins + 63) // 64 * 64)
        self._buf = {}
        self.refresh()

    # ---- weights -------------------------------------------------------------------------------
    def refresh(self):
        for c in [self.reduce, self.aspp_out, self.dcn_offset, self.head] + self.branches + \
                [c for pair in self.blocks for c in pair]:
            c.refresh()
        f = lambda t: t.detach().float().contiguous()
        net, a, dc = self.net, self.aspp, self.dcn
        if not hasattr(self, 's5'):                          # frozen BatchNorm folds: once
            self.bn_scale, self.bn_shift = fold_bn(net.bn)
            self.s5, self.b5 = fold_bn(a.global_avg_pool[2])
            self.s1, _ = fold_bn(a.bn1)
            cg, k = self.C // self.groups, self.k
            self.dcn_wf = [torch.empty(cg, k * k * cg, 1, 1, dtype=torch.bfloat16, device=self.device).view(cg, 1, 1, k * k * cg)
                           for _ in range(self.groups)]
            self.dcn_wb = [torch.empty(k * k * cg, 1, 1, cg, dtype=torch.bfloat16, device=self.device)
                           for _ in range(self.groups)]
        self.gap_ws = f(a.global_avg_pool[1].weight.flatten(1) * self.s5[:, None])
        self.w5s = f(a.conv1.weight.detach()[:, 4 * self.mid:].flatten(1) * self.s1[:, None])
        cg, k = self.C // self.groups, self.k
        lib, wd = _lib.load(), dc.weight.detach()
        for g in range(self.groups):
            # group g as a 1x1 layer over K = (tap, channel): [co][tap][ci] forward, [tap][ci][co] data gradient
            _lib.check(lib.dhd_pack_conv_weights(ctypes.c_void_p(wd.data_ptr() + g * cg * cg * k * k * 4), cg, cg, k * k,
                                                 0, cg, None, _p(self.dcn_wf[g]), cg, _p(self.dcn_wb[g]), cg, 1,
                                                 _stream()), 'pack_conv_weights(dcn)')

    def _act(self, name, N, H, W, C):
        key = (name, N, H, W, C)
        if key not in self._buf:
            self._buf[key] = D.Act.empty(N, H, W, C, 1, self.device)
        return self._buf[key]

    def _f32(self, name, *shape):
        key = (name,) + shape
        if key not in self._buf:
            self._buf[key] = torch.empty(*shape, device=self.device)
        return self._buf[key]

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, x, mlp_input):
        """x: Act (B*N, C_in, fH, fW); mlp_input (B, N, 27).  Returns softmax height (B*N, H, fH, fW) fp32."""
        net, lin = self.net, self._linear
        N, H, W, C = x.N, x.H, x.W, self.C
        f = lambda t: t.detach().float().contiguous()
        m_in = mlp_input.reshape(-1, mlp_input.shape[-1]).contiguous().float()
        mlp, se = net.depth_mlp, net.depth_se
        h1 = lin(m_in, f(mlp.fc1.weight), f(mlp.fc1.bias), 'relu', self.bn_scale, self.bn_shift)
        h2 = lin(h1, f(mlp.fc2.weight), f(mlp.fc2.bias))
        h3 = lin(h2, f(se.conv_reduce.weight.flatten(1)), f(se.conv_reduce.bias), 'relu')
        gate = lin(h3, f(se.conv_expand.weight.flatten(1)), f(se.conv_expand.bias), 'sigmoid')
        nhwc = D.nhwc_strides(C, H, W)
        h = self._act('h0', N, H, W, C)
        h32 = self._f32('h0_32', N, H, W, C)
        self.reduce.forward(x, [dict(act='relu', out_act=h, out_f32=(h32, nhwc))], img_gate=gate)
        hs, ts = [h], []
        for bi, (c1, c2) in enumerate(self.blocks):
            t = self._act('t%d' % bi, N, H, W, C)
            c1.forward(h, [dict(act='relu', out_act=t)])
            hn, hn32 = self._act('h%d' % (bi + 1), N, H, W, C), self._f32('h%d_32' % (bi + 1), N, H, W, C)
            c2.forward(t, [dict(act='relu', out_act=hn, out_f32=(hn32, nhwc))], residual=(h32, nhwc[:3]))
            h, h32 = hn, hn32
            hs.append(h)
            ts.append(t)
        mid = self.mid
        cat = self._act('cat', N, H, W, 4 * mid)
        for b, conv in enumerate(self.branches):
            conv.forward(h, [dict(act='relu', out_act=cat.slice(b * mid, (b + 1) * mid))])
        meanh = self._mean(h)
        x5 = lin(meanh, self.gap_ws, self.b5.to(self.device), 'relu')
        ib = lin(x5, self.w5s)
        ha = self._act('ha', N, H, W, C)
        self.aspp_out.forward(cat, [dict(act='relu', out_act=ha)], img_bias=ib)
        k, g, cg = self.k, self.groups, C // self.groups
        off = self._f32('off', N, H, W, self.noff)
        self.dcn_offset.forward(ha, [dict(out_f32=(off, D.nhwc_strides(self.noff, H, W)))])
        col = self._act('col', N, H, W, k * k * C)
        _lib.check(_lib.load().dhd_dcn_im2col(_p(ha.data), ha.ld, ha.coff, ha.part_stride, ha.parts, C, N, H, W,
                                              _p(off), self.noff, k, self.pad, self.dil, g, _p(col.data), col.ld,
                                              col.part_stride, col.parts, _stream()), 'dcn_im2col')
        out = self._act('dcn_out', N, H, W, C)
        for gi in range(g):
            D.conv2d(col.slice(gi * k * k * cg, (gi + 1) * k * k * cg), self.dcn_wf[gi], cg, precision='bf16',
                     segs=[dict(out_act=out.slice(gi * cg, (gi + 1) * cg))])
        height = torch.empty(N, self.H_bins, H, W, device=self.device)
        self.head.forward(out, [dict(act='softmax', out_f32=(height, D.nchw_strides(self.H_bins, H, W)))])
        self.saved = dict(x=x, m_in=m_in, h1=h1, h2=h2, h3=h3, gate=gate, hs=hs, ts=ts, cat=cat, meanh=meanh, x5=x5,
                          ha=ha, off=off, col=col, out=out, height=height)
        return height

    # ---- loss ----------------------------------------------------------------------------------
    def loss(self, label, fg):
        """label (npix,) int32 GT height bin (-1: none), fg (npix,) uint8/bool: MGHS.get_height_loss on
        binned labels.  Returns a 1-element device tensor; keeps d loss / d logits for backward()."""
        sv = self.saved
        h = sv['height']
        N, Hb, H, W = h.shape
        fg = fg.to(torch.uint8).contiguous()
        nfg = fg.sum().float().reshape(1)
        dz = self._act('dz', N, H, W, self.head.cout_pad)
        res = torch.empty(1, device=self.device)
        _lib.check(_lib.load().dhd_height_loss(_p(h), _p(label.int().contiguous()), _p(fg), N, Hb, H * W,
                                               self.loss_weight, _p(nfg), _p(res), _p(dz.data), dz.ld, _stream()),
                   'height_loss')
        self.dz = dz
        return res

    # ---- backward ------------------------------------------------------------------------------
    def backward(self, want_dx=False):
        sv, lin, lib = self.saved, self._linear, _lib.load()
        x, hs, ts, cat, ha, off, col, out = sv['x'], sv['hs'], sv['ts'], sv['cat'], sv['ha'], sv['off'], sv['col'], sv['out']
        N, H, W, C = x.N, x.H, x.W, self.C
        HW = H * W
        k, g, cg, mid = self.k, self.groups, C // self.groups, self.mid
        nhwc = D.nhwc_strides(C, H, W)
        dz = self.dz
        _, sums = act_bwd(dz, None, None, want_sums=True)
        dout = self._act('d_out', N, H, W, C)
        self.head.backward(out, dz, [dict(out_act=dout)], bias_sums=sums[0])
        # ---- DCN: grouped GEMM backward, then the sampling backward
        dcol = self._act('d_col', N, H, W, k * k * C)
        wgrad = _ensure_grad(self.dcn.weight)
        for gi in range(g):
            cs, ds = col.slice(gi * k * k * cg, (gi + 1) * k * k * cg), dout.slice(gi * cg, (gi + 1) * cg)
            dw = D.conv2d_wgrad(cs, ds, cg)                                   # (cg, 1, k*k*cg), K = (tap, c)
            wgrad[gi * cg:(gi + 1) * cg].add_(dw.view(cg, k, k, cg).permute(0, 3, 1, 2))
            D.conv2d(ds, self.dcn_wb[gi], k * k * cg, precision='bf16',
                     segs=[dict(out_act=dcol.slice(gi * k * k * cg, (gi + 1) * k * k * cg))])
        dxs = self._f32('d_sample', N, H, W, C)
        doff = self._f32('d_off', N, H, W, self.noff)
        _lib.check(lib.dhd_dcn_col2im_bwd(_p(dcol.data), dcol.ld, _p(ha.data), ha.ld, ha.coff, C, N, H, W, _p(off),
                                          self.noff, k, self.pad, self.dil, g, _p(dxs), _p(doff), _stream()),
                   'dcn_col2im_bwd')
        doff_a = self._act('d_off_a', N, H, W, self.dcn_offset.cout_pad)
        doff_a.data.zero_()
        doff_a.data[..., :self.noff] = doff.to(torch.bfloat16)
        _, sums = act_bwd(doff_a, None, None, want_sums=True)
        dha = self._act('d_ha', N, H, W, C)
        self.dcn_offset.backward(ha, doff_a, [dict(out_act=dha)], bias_sums=sums[0], residual=(dxs, nhwc[:3]))
        # ---- ASPP
        act_bwd(dha, ha, 'relu')
        dib = self._mean(dha) * float(HW)                                       # (N, C): sum over pixels
        dcat = self._act('d_cat', N, H, W, 4 * mid)
        self.aspp_out.backward(cat, dha, [dict(out_act=dcat)])
        x5, meanh = sv['x5'], sv['meanh']
        s1, s5 = self.s1.to(self.device), self.s5.to(self.device)
        a = self.aspp
        _ensure_grad(a.conv1.weight)[:, 4 * mid:].add_((lin(dib.t().contiguous(), x5.t().contiguous()) *
                                                        s1[:, None]).view(C, mid, 1, 1))
        dg = (lin(dib, self.w5s.t().contiguous()) * (x5 > 0).float()).contiguous()
        _acc(a.global_avg_pool[1].weight, lin(dg.t().contiguous(), meanh.t().contiguous()) * s5[:, None])
        dmean = lin(dg, self.gap_ws.t().contiguous()) * (1.0 / HW)
        act_bwd(dcat, cat, 'relu')
        hl = hs[-1]
        A, B = self._f32('acc_a', N, H, W, C), self._f32('acc_b', N, H, W, C)
        src = None
        for b, conv in enumerate(self.branches):
            dst = A if b % 2 == 0 else B
            kw = {} if src is None else dict(residual=(src, nhwc[:3]))
            conv.backward(hl, dcat.slice(b * mid, (b + 1) * mid), [dict(out_f32=(dst, nhwc))], **kw)
            src = dst
        dh = self._act('d_h', N, H, W, C)
        _lib.check(lib.dhd_add_rowvec(_p(src), _p(dmean.contiguous()), N, HW, C, _p(dh.data), dh.ld, dh.coff,
                                      _stream()), 'add_rowvec')
        act_bwd(dh, hl, 'relu')
        # ---- BasicBlocks, last to first
        for bi in range(len(self.blocks) - 1, -1, -1):
            c1, c2 = self.blocks[bi]
            t, hin = ts[bi], hs[bi]
            dt = self._act('d_t', N, H, W, C)
            c2.backward(t, dh, [dict(out_act=dt)])
            act_bwd(dt, t, 'relu')
            dhin = self._act('d_hin%d' % (bi & 1), N, H, W, C)
            c1.backward(hin, dt, [dict(out_act=dhin)])
            act_bwd(dhin, hin if bi > 0 else None, 'relu' if bi > 0 else None, add=dh)     # + identity path
            dh = dhin
        # ---- reduce conv + SE gate
        gate = sv['gate']
        dpre = self._act('d_pre', N, H, W, C)
        gsum = self._f32('gsum', N, C)
        ws = _workspace(self.device, lib.dhd_sfa_gate_bwd_workspace_bytes(N, HW, C))
        _lib.check(lib.dhd_se_gate_bwd(_p(dh.data), dh.ld, dh.coff, _p(hs[0].data), hs[0].ld, hs[0].coff, C, N, HW,
                                       _p(gate), _p(dpre.data), dpre.ld, dpre.coff, _p(gsum), _p(ws), _stream()),
                   'se_gate_bwd')
        _, sums = act_bwd(dpre, None, None, want_sums=True)
        dx = self._act('d_x', N, H, W, x.C) if want_dx else None
        self.reduce.backward(x, dpre, [dict(out_act=dx)] if want_dx else None, bias_sums=sums[0])
        # ---- camera-aware gate MLP (B*N rows)
        net = self.net
        mlp, se = net.depth_mlp, net.depth_se
        f = lambda t: t.detach().float().contiguous()
        h1, h2, h3, m_in = sv['h1'], sv['h2'], sv['h3'], sv['m_in']
        m_bn = m_in * self.bn_scale.to(self.device) + self.bn_shift.to(self.device)
        dze = (gsum * gate * (1.0 - gate)).contiguous()
        _acc(se.conv_expand.weight, lin(dze.t().contiguous(), h3.t().contiguous()))
        _acc(se.conv_expand.bias, dze.sum(0))
        dzr = (lin(dze, f(se.conv_expand.weight.flatten(1)).t().contiguous()) * (h3 > 0).float()).contiguous()
        _acc(se.conv_reduce.weight, lin(dzr.t().contiguous(), h2.t().contiguous()))
        _acc(se.conv_reduce.bias, dzr.sum(0))
        dh2 = lin(dzr, f(se.conv_reduce.weight.flatten(1)).t().contiguous())
        _acc(mlp.fc2.weight, lin(dh2.t().contiguous(), h1.t().contiguous()))
        _acc(mlp.fc2.bias, dh2.sum(0))
        dz1 = (lin(dh2, f(mlp.fc2.weight).t().contiguous()) * (h1 > 0).float()).contiguous()
        _acc(mlp.fc1.weight, lin(dz1.t().contiguous(), m_bn.t().contiguous()))
        _acc(mlp.fc1.bias, dz1.sum(0))
        return dx
