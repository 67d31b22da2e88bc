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
# How the BatchNorm after a convolution behaves on the training path: 'frozen' (running statistics and affine
# parameters fixed, folded into the convolution epilogue) or 'batch' (torch's training mode: batch statistics,
# trainable gamma / beta, running statistics updated).  Read when a trainer is constructed.
BN_MODE = 'frozen'
# batch-statistics BatchNorm: sums from the convolution epilogue (True) or from a separate pass over the output (A/B)
import os as _os
FUSED_BN_STATS = _os.environ.get('DHD_FUSED_BN_STATS', '1') != '0'
# the occupancy head's Linear + Softplus + Linear in training: one fused launch (True) or two convolution launches (A/B)
FUSED_TAIL = _os.environ.get('DHD_TRAIN_FUSED_TAIL', '1') != '0'
# bilinear up-sampling backward (FPN_LSS): deterministic gather kernel writing bf16 (True) or the atomic scatter into fp32 (A/B)
UPSAMPLE_BWD_GATHER = _os.environ.get('DHD_UPSAMPLE_BWD_GATHER', '1') != '0'
# SFA backward: dL/dx collected in bf16 activations (True) or in fp32 tensors (A/B)
SFA_BF16_DX = _os.environ.get('DHD_SFA_BF16_DX', '1') != '0'


_BN_COUNTERS = []


def flush_bn_counters():
    """`num_batches_tracked += 1` for every BatchNorm that ran on batch statistics since the last call (torch bumps the
    counter in each training-mode forward; here the bumps of a whole step are ONE multi-tensor launch)."""
    global _BN_COUNTERS
    if _BN_COUNTERS:
        cs, _BN_COUNTERS = _BN_COUNTERS, []
        torch._foreach_add_(cs, 1)


def set_bn_mode(mode):
    global BN_MODE
    if mode not in ('frozen', 'batch'):
        raise ValueError(mode)
    BN_MODE = mode
_WS = {}


def _workspace(dev, nbytes):
    key = (dev, torch.cuda.current_stream().cuda_stream, 'act')
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WS[key] = ws
    return ws


def act_bwd(dy, y, act, out=None, want_sums=False, add=None, write=True):
    """dz = (dy [+ add]) * act'(y) (Acts, part 0).  out=None -> in place on dy.  Returns (dz Act, sums)
    with sums (2, C) fp32 = [sum_rows dz, sum_rows dz*y] when want_sums."""
    lib = _lib.load()
    C = dy.C
    in_place = out is None
    out = dy if out is None else out
    if not write or (in_place and act in (None, 'none') and add is None):
        out = None                                # reductions only (an in-place identity pass would rewrite dy unchanged)
    sums = ws = None
    if want_sums:
        sums = torch.empty(2, C, device=dy.data.device)
        ws = _workspace(dy.data.device, lib.dhd_act_bwd_workspace_bytes(C))
    yy = y if y is not None else dy
    _lib.check(lib.dhd_act_bwd(_p(dy.data), dy.ld, dy.coff, _p(yy.data), yy.ld, yy.coff, dy.N * dy.H * dy.W, C,
                               ACT_ID[act], _p(out.data) if out is not None else None, out.ld if out is not None else 0,
                               out.coff if out is not None else 0, _p(sums), _p(ws),
                               _p(add.data) if add is not None else None, add.ld if add is not None else 0,
                               add.coff if add is not None else 0, _stream()), 'act_bwd')
    return (out if out is not None else dy), sums


class PackDesc(ctypes.Structure):
    """struct dhd_pack_desc (include/dhd_b200.h)."""
    _fields_ = [('w', ctypes.c_void_p), ('scale', ctypes.c_void_p), ('fwd', ctypes.c_void_p), ('bwd', ctypes.c_void_p),
                ('Cout', ctypes.c_int32), ('cin_total', ctypes.c_int32), ('taps', ctypes.c_int32), ('col_lo', ctypes.c_int32),
                ('Cin', ctypes.c_int32), ('cin_pad', ctypes.c_int32), ('cout_pad', ctypes.c_int32), ('bwd_mode', ctypes.c_int32)]


_PACK_BATCH = None
PACK_MAX_BATCH = 32


class batched_repack:
    """Context manager: _TrainConv.refresh() calls inside it queue their weight re-pack and the queue is flushed
    PACK_MAX_BATCH layers per launch (dhd_pack_conv_weights_batch) on exit."""

    def __enter__(self):
        global _PACK_BATCH
        self.prev, _PACK_BATCH = _PACK_BATCH, []
        return self

    def __exit__(self, *exc):
        global _PACK_BATCH
        queue, _PACK_BATCH = _PACK_BATCH, self.prev
        if exc[0] is not None:
            return False
        lib = _lib.load()
        for i in range(0, len(queue), PACK_MAX_BATCH):
            chunk = queue[i:i + PACK_MAX_BATCH]
            arr = (PackDesc * len(chunk))()
            for d, (w, scale, fwd, bwd, Cout, cin_total, taps, col_lo, Cin, cin_pad, cout_pad, mode) in zip(arr, chunk):
                d.w, d.scale = w.data_ptr(), (scale.data_ptr() if scale is not None else None)
                d.fwd, d.bwd = fwd.data_ptr(), bwd.data_ptr()
                d.Cout, d.cin_total, d.taps, d.col_lo, d.Cin = Cout, cin_total, taps, col_lo, Cin
                d.cin_pad, d.cout_pad, d.bwd_mode = cin_pad, cout_pad, mode
            _lib.check(lib.dhd_pack_conv_weights_batch(arr, len(chunk), _stream()), 'pack_conv_weights_batch')
        return False


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

    def __init__(self, weight, bias, bn, ksize, dilation=1, cin_pad=None, cout_pad=None, cols=None, stride=1):
        """cols=(lo, hi): the layer uses input-channel columns [lo, hi) of `weight` only (a branch of a
        concatenation whose other columns are handled elsewhere)."""
        self.weight, self.bias_p, self.bn = weight, bias, bn
        self.batch_bn = bn is not None and BN_MODE == 'batch'
        if self.batch_bn and cout_pad not in (None, weight.shape[0]) and weight.shape[0] % 8 != 0:
            raise NotImplementedError('batch-statistics BatchNorm on a padded output needs Cout % 8 == 0')
        self.ksize, self.dilation, self.cols, self.stride = ksize, dilation, cols, stride
        if stride not in (1, 2) or (stride == 2 and (ksize != 3 or dilation != 1)):
            raise NotImplementedError('stride-2 layers: 3x3, pad 1')
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
            if self.bn is not None and not self.batch_bn:
                s, b = fold_bn(self.bn, None)
                self.scale, self._bn_bias = s.to(dev), b.to(dev)
            else:
                self.scale = None                  # batch statistics: nothing to fold, BN runs as its own pass
            self.w_fwd = torch.empty(self.Cout, taps, 1, self.cin_pad, dtype=torch.bfloat16, device=dev)
            # data-gradient weight [ci][tap][co]: rows beyond Cin (zero-padded input channels) stay zero
            self.w_bwd = torch.zeros(self.cin_pad, taps, 1, self.cout_pad, dtype=torch.bfloat16, device=dev)
        if self.bn is None or self.batch_bn:
            self.bias = self.bias_p.detach() if self.bias_p is not None else None     # shares the parameter's storage
        else:
            self.bias = self._bn_bias if self.bias_p is None else self._bn_bias + self.bias_p.detach() * self.scale
        col_lo = 0 if self.cols is None else self.cols[0]
        if _PACK_BATCH is not None:
            _PACK_BATCH.append((w.detach(), self.scale, self.w_fwd, self.w_bwd, self.Cout, cin_total, taps, col_lo,
                                self.Cin, self.cin_pad, self.cout_pad, 0))
        else:
            _lib.check(_lib.load().dhd_pack_conv_weights(_p(w.detach()), self.Cout, cin_total, taps, col_lo, self.Cin,
                                                         _p(self.scale), _p(self.w_fwd), self.cin_pad, _p(self.w_bwd),
                                                         self.cout_pad, 0, _stream()), 'pack_conv_weights')
        if self.stride == 2:
            # data gradient of a stride-2 3x3 / pad-1 layer, one stride-1 convolution per output parity phase
            # (py, px): dx[2u+py, 2v+px] = sum over the taps k with k = p+1 (mod 2) of dy[u + (p+1-k)/2, ...] w[k]
            wd = w.detach().float()
            if self.scale is not None:
                wd = wd * self.scale.view(-1, 1, 1, 1)
            cand = {0: [(1, 0)], 1: [(2, 0), (0, 1)]}                  # parity -> [(kernel index, dy offset)], (.., 0) first
            self.phases = []
            for py in range(2):
                for px in range(2):
                    offs, ws = [], []
                    for ky, oy in cand[py]:
                        for kx, ox in cand[px]:
                            offs.append((oy, ox))
                            ws.append(wd[:, :, ky, kx].t())           # (Cin, Cout)
                    wp = torch.stack(ws, dim=1)                        # (Cin, taps, Cout)
                    if self.cout_pad != self.Cout or self.cin_pad != wp.shape[0]:      # zero columns / rows of the padding
                        wp = torch.nn.functional.pad(wp, (0, self.cout_pad - self.Cout, 0, 0, 0, self.cin_pad - wp.shape[0]))
                    self.phases.append((py, px, offs, wp.unsqueeze(2).to(torch.bfloat16).contiguous()))

    def forward(self, x, segs, **kw):
        if self.batch_bn:
            return self._forward_batch_bn(x, segs, **kw)
        return D.conv2d(x, self.w_fwd, self.Cout, ksize=self.ksize, dilation=self.dilation, precision='bf16',
                        scale=self.scale, bias=self.bias, segs=segs, stride=self.stride, **kw)

    # ---- BatchNorm with batch statistics: conv -> raw (bf16), statistics, normalise + activation as a second pass
    def _forward_batch_bn(self, x, segs, img_bias=None, img_gate=None, residual=None, residual_act=None):
        if len(segs) != 1:
            raise NotImplementedError('batch-statistics BatchNorm: one output segment')
        seg, bn, lib = segs[0], self.bn, _lib.load()
        oH, oW = (x.H, x.W) if self.stride == 1 else ((x.H + 1) // 2, (x.W + 1) // 2)
        key = (x.N, oH, oW)
        if getattr(self, '_raw_key', None) != key:
            self._raw = D.Act.empty(x.N, oH, oW, self.Cout, 1, x.data.device)
            self._dyraw = D.Act.empty(x.N, oH, oW, self.cout_pad, 1, x.data.device)
            if self.cout_pad != self.Cout:
                self._dyraw.data.zero_()
            self._raw_key = key
        raw = self._raw
        if FUSED_BN_STATS and self.Cout % 2 == 0:
            # [sum raw, sum raw^2] per channel come out of the convolution's epilogue (per-tile partials of the staged
            # bf16 tile + a fixed-order finish): the separate read of `raw` is gone
            part, prows = D.conv2d(x, self.w_fwd, self.Cout, ksize=self.ksize, dilation=self.dilation, precision='bf16',
                                   bias=self.bias, img_bias=img_bias, segs=[dict(out_act=raw)], stride=self.stride,
                                   stats='partial')
            sums = None
        else:
            D.conv2d(x, self.w_fwd, self.Cout, ksize=self.ksize, dilation=self.dilation, precision='bf16', bias=self.bias,
                     img_bias=img_bias, segs=[dict(out_act=raw)], stride=self.stride)
            _, sums = act_bwd(raw, raw, None, want_sums=True, write=False)            # [sum raw, sum raw^2]
        M = float(x.N * oH * oW)
        if getattr(self, '_bn_buf', None) is None:
            self._bn_buf = torch.empty(7, self.Cout, device=x.data.device)       # scale, shift, mean, invstd, k1, k2, k3
        bb = self._bn_buf
        track = bn.track_running_stats
        mom = bn.momentum if bn.momentum is not None else 0.1
        if sums is None:                          # finish of the epilogue's partial sums + coefficient formulas: one launch
            _lib.check(lib.dhd_bn_fwd_coeffs_partial(_p(part), prows, self.Cout, M, _p(bn.weight.detach()),
                                                     _p(bn.bias.detach()), bn.eps, mom,
                                                     _p(bn.running_mean) if track else None,
                                                     _p(bn.running_var) if track else None,
                                                     _p(bb[0]), _p(bb[1]), _p(bb[2]), _p(bb[3]), _stream()), 'bn_fwd_coeffs_partial')
        else:
            _lib.check(lib.dhd_bn_fwd_coeffs(_p(sums), self.Cout, M, _p(bn.weight.detach()), _p(bn.bias.detach()), bn.eps,
                                             mom, _p(bn.running_mean) if track else None, _p(bn.running_var) if track else None,
                                             _p(bb[0]), _p(bb[1]), _p(bb[2]), _p(bb[3]), _stream()), 'bn_fwd_coeffs')
        scale, shift = bb[0], bb[1]
        self._bn_saved = M
        if track and getattr(bn, 'num_batches_tracked', None) is not None:
            _BN_COUNTERS.append(bn.num_batches_tracked)
            if len(_BN_COUNTERS) >= 1024:            # a caller that never flushes (direct use of a trainer)
                flush_bn_counters()
        ob, of, f_ld = seg.get('out_act'), None, 0
        if seg.get('out_f32') is not None:
            of, st = seg['out_f32']
            if st[3] != 1 or st[2] < self.Cout:
                raise NotImplementedError('batch-statistics BatchNorm: fp32 output must be NHWC rows')
            f_ld = st[2]
        if residual_act is not None:              # bf16 identity path (image backbone): act(scale*raw + shift + identity)
            ra = residual_act
            if residual is not None or img_gate is not None or of is not None or ob is None or \
                    (ra.N, ra.H, ra.W) != (x.N, oH, oW) or ra.C < self.Cout:
                raise NotImplementedError('batch-statistics BatchNorm with a bf16 identity: one bf16 output, no gate')
            _lib.check(lib.dhd_bn_apply_res16(_p(raw.data), raw.ld, raw.coff, x.N * oH * oW, self.Cout, _p(scale), _p(shift),
                                              ACT_ID[seg.get('act')], _p(ra.data), ra.ld, ra.coff, _p(ob.data), ob.ld,
                                              ob.coff, _stream()), 'bn_apply_res16')
            return
        res, res_ld = (None, 0) if residual is None else (residual[0], residual[1][2])
        _lib.check(lib.dhd_bn_apply(_p(raw.data), raw.ld, raw.coff, x.N * oH * oW, self.Cout, _p(scale), _p(shift),
                                    ACT_ID[seg.get('act')], _p(res), res_ld, _p(img_gate), oH * oW,
                                    _p(ob.data) if ob is not None else None, ob.ld if ob is not None else 0,
                                    ob.coff if ob is not None else 0, _p(of), f_ld, _stream()), 'bn_apply')

    def _backward_batch_bn(self, dy):
        """dy: gradient at the BatchNorm output (pre-activation).  Accumulates d gamma / d beta and returns the gradient
        at the convolution output (a scratch activation: dy itself may feed other layers and is left untouched)."""
        bn, raw, bb, M = self.bn, self._raw, self._bn_buf, self._bn_saved
        train_affine = bn.weight.requires_grad
        lib = _lib.load()
        # [sum dz, sum dz * raw] per channel (one pass over dy and raw), then finish + k1 / k2 / k3 + d gamma / d beta
        ws = _workspace(dy.data.device, lib.dhd_act_bwd_workspace_bytes(self.Cout))
        _lib.check(lib.dhd_bn_bwd_sums_coeffs(_p(dy.data), dy.ld, dy.coff, _p(raw.data), raw.ld, raw.coff,
                                              raw.N * raw.H * raw.W, self.Cout, _p(ws), M, _p(bb[2]), _p(bb[3]),
                                              _p(bn.weight.detach()), _p(bb[4]), _p(bb[5]), _p(bb[6]),
                                              _p(_ensure_grad(bn.weight)) if train_affine else None,
                                              _p(_ensure_grad(bn.bias)) if train_affine else None, _stream()),
                   'bn_bwd_sums_coeffs')
        out = self._dyraw
        _lib.check(_lib.load().dhd_affine_combine(_p(dy.data), dy.ld, dy.coff, _p(raw.data), raw.ld, raw.coff,
                                                  raw.N * raw.H * raw.W, self.Cout, _p(bb[4]), _p(bb[5]), _p(bb[6]),
                                                  _p(out.data), out.ld, out.coff, _stream()), 'affine_combine')
        return out

    def backward(self, x, dy, dx_segs=None, bias_sums=None, **kw):
        """dy: Act = gradient w.r.t. this layer's pre-activation output (channels >= Cout zero).
        Accumulates weight / bias gradients; writes dx through `dx_segs` (conv2d segs) if given."""
        if self.batch_bn:
            dy = self._backward_batch_bn(dy)
            bias_sums = None                       # a conv bias under a batch-statistics BN has zero gradient
            if self.bias_p is not None and self.bias_p.requires_grad:
                _ensure_grad(self.bias_p)          # ... which torch reports as zeros, not as None
        # the reduce pass of the weight-gradient kernel adds straight into param.grad (torch layout, its column window)
        gw = _ensure_grad(self.weight)
        if gw.is_contiguous() and gw.dtype == torch.float32:
            D.conv2d_wgrad(x, dy, self.Cout, ksize=self.ksize, dilation=self.dilation, scale=self.scale,
                           stride=self.stride, grad=(gw, 0 if self.cols is None else self.cols[0], self.Cin))
        else:
            dw = D.conv2d_wgrad(x, dy, self.Cout, ksize=self.ksize, dilation=self.dilation, scale=self.scale,
                                stride=self.stride)
            if x.C != self.Cin:
                dw = dw[:, :, :self.Cin]
            g = D.weight_grad_to_torch(dw.contiguous(), self.ksize)
            if self.cols is not None:
                gw[:, self.cols[0]:self.cols[1]].add_(g)
            else:
                _acc(self.weight, g if self.weight.dim() == 4 else g[:, :, 0, 0])
        if self.bias_p is not None and bias_sums is not None:      # a conv bias under a frozen BN sees the BN scale
            _acc(self.bias_p, bias_sums[:self.Cout] if self.bn is None else bias_sums[:self.Cout] * self.scale)
        if dx_segs is not None and self.stride == 2:
            if x.H % 2 or x.W % 2 or kw or len(dx_segs) != 1 or dx_segs[0].get('out_act') is None:
                raise NotImplementedError('stride-2 data gradient: even input size, one bf16 output')
            out = dx_segs[0]['out_act']
            ld = out.ld
            for py, px, offs, wp in self.phases:
                D.conv2d(dy, wp, self.cin_pad, precision='bf16', taps=offs,
                         segs=[dict(out_act=out, out_view=(x.H * x.W * ld, 2 * x.W * ld, 2 * ld, (py * x.W + px) * ld))])
        elif dx_segs is not None:
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
        Co = self.nout
        out = torch.empty(N, W, H, Co, device=self.device)
        if FUSED_TAIL and self.conv.Cout == 256 and self.fc0.Cout == 512 and self.Dz == 16 and self.ncls == 18 and \
                self.fc0.cin_pad == 256 and self.fc2.cin_pad == 512:
            # Linear + Softplus + Linear as the back-to-back GEMM kernel of the inference path (dhd_predictor_tail); the
            # hidden layer is written once for the backward instead of written by one launch and re-read by the next
            D.predictor_tail(t, self.fc0.w_fwd, self.fc0.bias, self.fc2.w_fwd, self.fc2.bias, self.Dz, self.ncls,
                             logits=out, hidden=u)
        else:
            self.fc0.forward(t, [dict(act='softplus', out_act=u)])
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
        lean = SFA_BF16_DX
        dpre2 = self._act('dpre2', N, H, W, C)
        S = self._f32('S', N, C)
        ws = _workspace(self.device, lib.dhd_sfa_gate_bwd_workspace_bytes(N, H * W, C))
        if lean:
            # dL/dx is collected in bf16 activations: the fuse blend's part (dxa), + the shortcut's data gradient through
            # the bf16 residual input of its epilogue (dxb), + the channel-gate blend's and the squeeze path's parts in
            # one final pass (dhd_sfa_dx_combine) -- instead of 2C fp32 values per pixel written, re-read, read-modified-
            # written and read again (~2 GB of traffic at DHD-S B=4)
            dxa = self._act('dxa16', N, H, W, 2 * C)
            dxb = self._act('dxb16', N, H, W, 2 * C)
            _lib.check(lib.dhd_sfa_gate_bwd_b16(0, _p(dfuse.data), dfuse.ld, dfuse.coff, _p(x.data), x.ld, x.coff, C, N, H * W,
                                                _p(a1), _p(a2), _p(dpre2.data), dpre2.ld, dpre2.coff, _p(dxa.data), dxa.ld,
                                                dxa.coff, _p(S), 0, _p(ws), _stream()), 'sfa_gate_bwd_b16(fuse)')
            self.short.backward(x, dout, [dict(out_act=dxb)], residual_act=dxa)
        else:
            dxa = self._f32('dxa', N, H, W, 2 * C)
            dxb = self._f32('dxb', N, H, W, 2 * C)
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
        if lean:
            _lib.check(lib.dhd_sfa_gate_bwd_b16(1, _p(du.data), du.ld, du.coff, _p(x.data), x.ld, x.coff, C, N, H * W,
                                                _p(a1), None, None, 0, 0, None, 0, 0, _p(S), 1, _p(ws), _stream()),
                       'sfa_gate_bwd_b16(u)')
        else:
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
        if lean:
            dx = dxb                              # in place: every element is read and written by the same thread
            _lib.check(lib.dhd_sfa_dx_combine(_p(dxb.data), dxb.ld, dxb.coff, _p(du.data), du.ld, du.coff, _p(a1),
                                              _p(ds.contiguous()), C, N, H * W, _p(dx.data), dx.ld, dx.coff, _stream()),
                       'sfa_dx_combine')
            return dx
        dx = self._act('dx', N, H, W, 2 * C)
        _lib.check(lib.dhd_add_rowvec(_p(dxb), _p(ds.contiguous()), N, H * W, 2 * C, _p(dx.data), dx.ld, dx.coff,
                                      _stream()), 'add_rowvec')
        return dx


def gt_downsample(gt, ds, lo, interval, nbins, label=None, valid=None, want_valid=False):
    """MGHS.get_downsampled_gt_depth / _height (lss_heightmap.py:625-701) as bin indices: gt (B, N, H, W) fp32 sparse
    map -> label (B*N*H/ds*W/ds,) int32 in [-1, nbins) (the index of the reference's one-hot row, -1 = all zero) and,
    with want_valid, the uint8 flag label >= 0.  depth: lo = d_min - d_step, interval = d_step; height: lo =
    height_range[0], interval = height_interval."""
    if not gt.is_cuda:
        raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
    gt = gt.float().contiguous()
    H, W = gt.shape[-2:]
    BN = gt.numel() // (H * W)
    n = BN * (H // ds) * (W // ds)
    if label is None:
        label = torch.empty(n, dtype=torch.int32, device=gt.device)
    if valid is None and want_valid:
        valid = torch.empty(n, dtype=torch.uint8, device=gt.device)
    _lib.check(_lib.load().dhd_gt_downsample(_p(gt), BN, H, W, int(ds), float(lo), float(interval), int(nbins), _p(label),
                                             _p(valid), _stream()), 'gt_downsample')
    return label, valid


def lidar_losses(vt, gt_depth, gt_height, depth=None, height=None):
    """MGHS.get_height_loss (lss_heightmap.py:595-622) and MGHS_Depth.get_depth_and_height_loss (859-897) on the CUDA
    kernels: the sparse gt maps are binned by dhd_gt_downsample (foreground = pixels whose depth return lands in a
    depth bin of vt.grid_config['depth'], the config the module holds at loss time), then one dhd_height_loss launch per
    distribution gives weight * sum_fg BCE / max(1, n_fg) and its gradient at the logits (through the softmax).
    depth / height: softmax probabilities (B*N, bins, fH, fW) fp32, either may be None.
    Returns {'depth': (loss (1,), dz Act), 'height': (loss, dz Act)} for the ones given."""
    if getattr(vt, 'sid', False):
        raise NotImplementedError('dhd_gt_downsample bins depth linearly; sid=True (log-spaced bins, lss_heightmap.py:646-653) '
                                  'is not used by any DHD config -- use the plugin torch form MGHS.get_downsampled_gt_depth')
    dc = vt.grid_config['depth']
    dlab, fg = gt_downsample(gt_depth, vt.downsample, dc[0] - dc[2], dc[2], vt.D, want_valid=True)
    hlab, _ = gt_downsample(gt_height, vt.downsample, vt.height_range[0], vt.height_interval, vt.H)
    nfg = fg.sum().float().reshape(1)
    out = {}
    for name, probs, lab, w in (('depth', depth, dlab, getattr(vt, 'loss_depth_weight', 0.0)),
                                ('height', height, hlab, vt.loss_height_weight)):
        if probs is None:
            continue
        if not probs.is_cuda or probs.dtype != torch.float32:
            raise RuntimeError('dhd_b200: expected fp32 CUDA probabilities (the hot path has no CPU fallback)')
        N, K, H, W = probs.shape
        dz = D.Act.empty(N, H, W, (K + 63) // 64 * 64, 1, probs.device)
        res = torch.empty(1, device=probs.device)
        _lib.check(_lib.load().dhd_height_loss(_p(probs.contiguous()), _p(lab), _p(fg), N, K, H * W, float(w), _p(nfg),
                                               _p(res), _p(dz.data), dz.ld, _stream()), 'height_loss')
        out[name] = (res, dz)
    return out


def dropout_(a, p, rng, salt=0):
    """In-place Dropout of an Act (bf16 part 0) with the counter-based mask of dhd_dropout."""
    _lib.check(_lib.load().dhd_dropout(_p(a.data), a.ld, a.coff, a.N * a.H * a.W, a.C, float(p), _p(rng), int(salt),
                                       _stream()), 'dropout')
    return a


def _ensure_grad(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


class HeightNetTrainer:
    """HeightNet (depthnet.py:418-487, 605-652, non-stereo) -- and, through `DepthNetTrainer`, the trunk of the
    camera-aware DepthNet; BatchNorm frozen or on batch statistics (set_bn_mode), the ASPP's Dropout on when
    `dropout` > 0: reduce conv + camera-aware SE gate, BasicBlocks (the first one optionally on a concatenated input
    with a 1x1 `downsample` identity path: stereo DepthNet), ASPP (global branch as a per-image bias), optional DCN
    (deformable im2col + grouped GEMM; off in DHD-M / DHD-L), 1x1 head + softmax; backward of all of it, fed by the
    height / depth loss (lss_heightmap.py:595-622, 859-897)."""

    def __init__(self, net, device='cuda', loss_weight=0.1, dropout=0.0, seed=0):
        """dropout: drop probability of the ASPP's nn.Dropout (depthnet.py:81; 0.5 in the reference's train mode,
        0 = eval behaviour).  The mask comes from a counter-based generator (dhd_dropout) keyed by (seed, step); the
        step counter lives on the device and is bumped at the end of backward()."""
        from .modules import linear_rows, mean_hw
        self._linear, self._mean = linear_rows, mean_hw
        self.net, self.device, self.loss_weight = net, device, float(loss_weight)
        self.dropout_p = float(dropout)
        self.rng = torch.tensor([int(seed), 0], dtype=torch.int64, device=device)
        self.C = net.reduce_conv[0].out_channels
        self.reduce = _TrainConv(net.reduce_conv[0].weight, net.reduce_conv[0].bias, net.reduce_conv[1], 3)
        layers = list(net.depth_conv)
        self.blocks, i = [], 0
        self.cat_channels = None
        while i < len(layers) and type(layers[i]).__name__ == 'BasicBlock':
            b = layers[i]
            if b.downsample is not None:
                # stereo DepthNet (depthnet.py:203-218): the first block reads cat(gated feature, cost_volumn_net output),
                # channels zero-padded to the 64-channel granule of the concatenation buffer, and carries a plain 1x1
                # convolution (with bias, no BatchNorm) on its identity path
                cin = b.conv1.weight.shape[1]
                self.cat_channels = (cin + 63) // 64 * 64
                self.blocks.append((_TrainConv(b.conv1.weight, None, b.bn1, 3, cin_pad=self.cat_channels),
                                    _TrainConv(b.conv2.weight, None, b.bn2, 3),
                                    _TrainConv(b.downsample.weight, b.downsample.bias, None, 1, cin_pad=self.cat_channels)))
            else:
                self.blocks.append((_TrainConv(b.conv1.weight, None, b.bn1, 3), _TrainConv(b.conv2.weight, None, b.bn2, 3), None))
            i += 1
        a = layers[i]
        if type(a).__name__ != 'ASPP':
            raise NotImplementedError('trunk without ASPP')
        self.aspp = a
        self.mid = mid = a.aspp1.atrous_conv.out_channels
        self.mid_pad = (mid + 63) // 64 * 64              # branch slices of the concatenation buffer start on 64-channel granules
        self.branches = [_TrainConv(b.atrous_conv.weight, None, b.bn, b.atrous_conv.kernel_size[0], b.atrous_conv.dilation[0],
                                    cout_pad=self.mid_pad)
                         for b in (a.aspp1, a.aspp2, a.aspp3, a.aspp4)]
        if self.mid_pad != mid:
            # aspp_mid_channels=96 (DHD-M / DHD-L): conv1's input columns are spread over the padded slices
            w = a.conv1.weight
            self._aspp_cols = torch.cat([torch.arange(b * self.mid_pad, b * self.mid_pad + mid) for b in range(4)])
            self.aspp_out = _TrainConvScatter(w, a.bn1, 4 * mid, 4 * self.mid_pad, self._aspp_cols)
        else:
            self.aspp_out = _TrainConv(a.conv1.weight, None, a.bn1, 1, cols=(0, 4 * mid))
        i += 1
        self.dcn = layers[i] if hasattr(layers[i], 'conv_offset') else None
        if self.dcn is not None:
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
    def _convs(self):
        cs = [self.reduce, self.aspp_out, self.head] + self.branches + [c for tr in self.blocks for c in tr if c is not None]
        if self.dcn is not None:
            cs.append(self.dcn_offset)
        return cs

    def refresh(self):
        for c in self._convs():
            c.refresh()
        f = lambda t: t.detach().float().contiguous()
        net, a, dc = self.net, self.aspp, self.dcn
        if not hasattr(self, 's5'):                          # frozen BatchNorm folds: once
            self.bn_scale, self.bn_shift = fold_bn(net.bn)
            self.s5, self.b5 = fold_bn(a.global_avg_pool[2])
            self.s1, _ = fold_bn(a.bn1)
            if self.aspp_out.batch_bn:                 # batch statistics: the global branch enters the conv sum unscaled
                self.s1 = torch.ones_like(self.s1)
            if dc is not None:
                cg, k = self.C // self.groups, self.k
                self.dcn_wf = [torch.empty(cg, k * k * cg, 1, 1, dtype=torch.bfloat16, device=self.device).view(cg, 1, 1, k * k * cg)
                               for _ in range(self.groups)]
                self.dcn_wb = [torch.empty(k * k * cg, 1, 1, cg, dtype=torch.bfloat16, device=self.device)
                               for _ in range(self.groups)]
        self.gap_ws = f(a.global_avg_pool[1].weight.flatten(1) * self.s5[:, None])
        self.w5s = f(a.conv1.weight.detach()[:, 4 * self.mid:].flatten(1) * self.s1[:, None])
        if dc is None:
            return
        cg, k = self.C // self.groups, self.k
        lib, wd = _lib.load(), dc.weight.detach()
        for g in range(self.groups):
            # group g as a 1x1 layer over K = (tap, channel): [co][tap][ci] forward, [tap][ci][co] data gradient
            _lib.check(lib.dhd_pack_conv_weights(ctypes.c_void_p(wd.data_ptr() + g * cg * cg * k * k * 4), cg, cg, k * k,
                                                 0, cg, None, _p(self.dcn_wf[g]), cg, _p(self.dcn_wb[g]), cg, 1,
                                                 _stream()), 'pack_conv_weights(dcn)')

    def _act(self, name, N, H, W, C, zero=False):
        key = (name, N, H, W, C)
        if key not in self._buf:
            self._buf[key] = D.Act.empty(N, H, W, C, 1, self.device)
            if zero:
                self._buf[key].data.zero_()
        return self._buf[key]

    def _f32(self, name, *shape):
        key = (name,) + shape
        if key not in self._buf:
            self._buf[key] = torch.empty(*shape, device=self.device)
        return self._buf[key]

    # ---- camera-aware gate: BatchNorm1d (frozen) -> Mlp -> SELayer -> sigmoid (depthnet.py:119-169, 624-629) -------
    def _gate_forward(self, m_in, mlp, se):
        lin = self._linear
        f = lambda t: t.detach().float().contiguous()
        h1 = lin(m_in, f(mlp.fc1.weight), f(mlp.fc1.bias), 'relu', self.bn_scale, self.bn_shift)
        h2 = lin(h1, f(mlp.fc2.weight), f(mlp.fc2.bias))
        h3 = lin(h2, f(se.conv_reduce.weight.flatten(1)), f(se.conv_reduce.bias), 'relu')
        gate = lin(h3, f(se.conv_expand.weight.flatten(1)), f(se.conv_expand.bias), 'sigmoid')
        return dict(h1=h1, h2=h2, h3=h3, gate=gate)

    def _gate_backward(self, gsum, gs, m_in, mlp, se):
        """gsum (N, C) = dL/d gate (per image and channel) -> gradients of the SELayer and the Mlp."""
        lin = self._linear
        f = lambda t: t.detach().float().contiguous()
        h1, h2, h3, gate = gs['h1'], gs['h2'], gs['h3'], gs['gate']
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

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, x, mlp_input):
        """x: Act (B*N, C_in, fH, fW); mlp_input (B, N, 27).  Returns softmax height (B*N, H, fH, fW) fp32."""
        net = self.net
        N, H, W, C = x.N, x.H, x.W, self.C
        m_in = mlp_input.reshape(-1, mlp_input.shape[-1]).contiguous().float()
        gs = self._gate_forward(m_in, net.depth_mlp, net.depth_se)
        nhwc = D.nhwc_strides(C, H, W)
        h = self._act('h0', N, H, W, C)
        h32 = self._f32('h0_32', N, H, W, C)
        self.reduce.forward(x, [dict(act='relu', out_act=h, out_f32=(h32, nhwc))], img_gate=gs['gate'])
        self.saved = dict(x=x, m_in=m_in, gs=gs, gate=gs['gate'])
        return self._trunk_forward(h, h32)

    def _trunk_forward(self, h, h32):
        """BasicBlocks -> ASPP (-> Dropout) -> [DCN] -> 1x1 head + softmax on the (gated / concatenated) feature map `h`
        (Act; h32 = its fp32 copy for the first residual, None when the first block has a `downsample` path)."""
        lin = self._linear
        N, H, W, C = h.N, h.H, h.W, self.C
        nhwc = D.nhwc_strides(C, H, W)
        hs, ts = [h], []
        for bi, (c1, c2, ds) in enumerate(self.blocks):
            t = self._act('t%d' % bi, N, H, W, C)
            if ds is not None:                      # identity path = 1x1 convolution of the concatenated input
                h32 = self._f32('idn%d' % bi, N, H, W, C)
                ds.forward(h, [dict(out_f32=(h32, nhwc))])
            c1.forward(h, [dict(act='relu', out_act=t)])
            hn, hn32 = self._act('h%d' % (bi + 1), N, H, W, C), self._f32('h%d_32' % (bi + 1), N, H, W, C)
            c2.forward(t, [dict(act='relu', out_act=hn, out_f32=(hn32, nhwc))], residual=(h32, nhwc[:3]))
            h, h32 = hn, hn32
            hs.append(h)
            ts.append(t)
        mid, mp = self.mid, self.mid_pad
        cat = self._act('cat', N, H, W, 4 * mp, zero=True)
        for b, conv in enumerate(self.branches):
            conv.forward(h, [dict(act='relu', out_act=cat.slice(b * mp, b * mp + mid))])
        meanh = self._mean(h)
        x5 = lin(meanh, self.gap_ws, self.b5.to(self.device), 'relu')
        ib = lin(x5, self.w5s)
        ha = self._act('ha', N, H, W, C)
        self.aspp_out.forward(cat, [dict(act='relu', out_act=ha)], img_bias=ib)
        hook, self.trunk_hook = getattr(self, 'trunk_hook', None), None
        if hook is not None:                      # one-shot: the pipeline forks the binning kernels here (late, L2-hot bins)
            hook()
        if self.dropout_p > 0.0:                     # in place: everything downstream (and the backward) sees the dropped map
            dropout_(ha, self.dropout_p, self.rng, salt=1)
        sv = self.saved
        sv.update(hs=hs, ts=ts, cat=cat, meanh=meanh, x5=x5, ha=ha)
        out = ha
        if self.dcn is not None:
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
            sv.update(off=off, col=col)
        height = torch.empty(N, self.H_bins, H, W, device=self.device)
        self.head.forward(out, [dict(act='softmax', out_f32=(height, D.nchw_strides(self.H_bins, H, W)))])
        sv.update(out=out, height=height)
        return height

    # ---- loss ----------------------------------------------------------------------------------
    def loss(self, label, fg):
        """label (npix,) int32 GT bin (-1: none), fg (npix,) uint8/bool: MGHS.get_height_loss (and the depth term of
        get_depth_and_height_loss) on binned labels.  Returns a 1-element device tensor; keeps d loss / d logits for
        backward()."""
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
    def _trunk_backward(self):
        """From self.dz (gradient at the head's logits) down to the trunk's input: returns dL/d hs[0] as an Act with the
        channel count of that input (C, or cat_channels for the stereo first block) -- no activation mask applied."""
        sv, lin, lib = self.saved, self._linear, _lib.load()
        hs, ts, cat, ha, out = sv['hs'], sv['ts'], sv['cat'], sv['ha'], sv['out']
        N, H, W, C = ha.N, ha.H, ha.W, self.C
        HW = H * W
        mid, mp = self.mid, self.mid_pad
        nhwc = D.nhwc_strides(C, H, W)
        dz = self.dz
        _, sums = act_bwd(dz, None, None, want_sums=True)
        dha = self._act('d_ha', N, H, W, C)
        if self.dcn is None:
            self.head.backward(out, dz, [dict(out_act=dha)], bias_sums=sums[0])
        else:
            off, col = sv['off'], sv['col']
            k, g, cg = self.k, self.groups, C // self.groups
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
            self.dcn_offset.backward(ha, doff_a, [dict(out_act=dha)], bias_sums=sums[0], residual=(dxs, nhwc[:3]))
        # ---- ASPP (Dropout backward = the same mask on the gradient; the ReLU test `y > 0` on the dropped map is
        # still right for every kept element and the dropped ones are already zero)
        if self.dropout_p > 0.0:
            dropout_(dha, self.dropout_p, self.rng, salt=1)
        act_bwd(dha, ha, 'relu')
        dcat = self._act('d_cat', N, H, W, 4 * mp)
        self.aspp_out.backward(cat, dha, [dict(out_act=dcat)])
        # gradient of the per-image bias the global branch adds to the conv sum: pixel sum of the gradient at the
        # convolution output (behind the BatchNorm backward when it runs on batch statistics)
        dib = self._mean(self.aspp_out._dyraw if self.aspp_out.batch_bn else dha) * float(HW)
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
            conv.backward(hl, dcat.slice(b * mp, (b + 1) * mp), [dict(out_f32=(dst, nhwc))], **kw)
            src = dst
        dh = self._act('d_h', N, H, W, C)
        _lib.check(lib.dhd_add_rowvec(_p(src), _p(dmean.contiguous()), N, HW, C, _p(dh.data), dh.ld, dh.coff,
                                      _stream()), 'add_rowvec')
        act_bwd(dh, hl, 'relu')
        # ---- BasicBlocks, last to first
        for bi in range(len(self.blocks) - 1, -1, -1):
            c1, c2, ds = self.blocks[bi]
            t, hin = ts[bi], hs[bi]
            dt = self._act('d_t', N, H, W, C)
            c2.backward(t, dh, [dict(out_act=dt)])
            act_bwd(dt, t, 'relu')
            dhin = self._act('d_hin%d' % (bi & 1), N, H, W, hin.C)
            c1.backward(hin, dt, [dict(out_act=dhin)])
            if ds is not None:                        # identity path = the 1x1 downsample convolution of the concatenated input
                _, sums = act_bwd(dh, None, None, want_sums=True)
                did = self._act('d_id', N, H, W, hin.C)
                ds.backward(hin, dh, [dict(out_act=did)], bias_sums=sums[0])
                act_bwd(dhin, None, None, add=did)
            else:
                act_bwd(dhin, hin if bi > 0 else None, 'relu' if bi > 0 else None, add=dh)     # + identity path
            dh = dhin
        return dh

    def backward(self, want_dx=False):
        sv, lib = self.saved, _lib.load()
        x = sv['x']
        N, H, W, C = x.N, x.H, x.W, self.C
        HW = H * W
        dh = self._trunk_backward()
        # ---- reduce conv + SE gate
        gate, h0 = sv['gate'], sv['hs'][0]
        dpre = self._act('d_pre', N, H, W, C)
        gsum = self._f32('gsum', N, C)
        ws = _workspace(self.device, lib.dhd_sfa_gate_bwd_workspace_bytes(N, HW, C))
        _lib.check(lib.dhd_se_gate_bwd(_p(dh.data), dh.ld, dh.coff, _p(h0.data), h0.ld, h0.coff, C, N, HW,
                                       _p(gate), _p(dpre.data), dpre.ld, dpre.coff, _p(gsum), _p(ws), _stream()),
                   'se_gate_bwd')
        _, sums = act_bwd(dpre, None, None, want_sums=True)
        dx = self._act('d_x', N, H, W, x.C) if want_dx else None
        self.reduce.backward(x, dpre, [dict(out_act=dx)] if want_dx else None, bias_sums=sums[0])
        # ---- camera-aware gate MLP (B*N rows)
        self._gate_backward(gsum, sv['gs'], sv['m_in'], self.net.depth_mlp, self.net.depth_se)
        if self.dropout_p > 0.0:
            self.rng[1:].add_(1)                   # next step, next mask (a device-side update: graph replays advance too)
        return dx


class _TrainConvScatter(_TrainConv):
    """A 1x1 convolution whose input channels sit at scattered positions of a wider, zero-padded buffer (the ASPP's
    conv1 when aspp_mid_channels is not a multiple of 64: the four branch slices start on 64-channel granules).  The
    master weight keeps the reference's (Cout, n_used + extra, 1, 1) shape; forward / data-gradient weights are packed
    from a scattered copy and the weight gradient is gathered back."""

    def __init__(self, weight, bn, n_used, n_padded, cols):
        self._full, self._cols, self._n_used, self._n_pad = weight, cols, n_used, n_padded
        self._wide = torch.zeros(weight.shape[0], n_padded, 1, 1, device=weight.device)
        super().__init__(self._wide, None, bn, 1)

    def refresh(self):
        self._wide.zero_()
        self._wide[:, self._cols.to(self._wide.device)] = self._full.detach()[:, :self._n_used].float()
        super().refresh()

    def backward(self, x, dy, dx_segs=None, bias_sums=None, **kw):
        self._wide.grad = None
        super().backward(x, dy, dx_segs, bias_sums, **kw)
        g = self._wide.grad
        _ensure_grad(self._full)[:, :self._n_used].add_(g[:, self._cols.to(g.device)])
        self._wide.grad = None


class DepthNetTrainer(HeightNetTrainer):
    """Camera-aware DepthNet of MGHS_Depth / MGHS_Stereo (depthnet.py:172-243, 362-415) in training: the reduce conv feeds
    TWO camera-gated branches -- context (SE gate -> 1x1 context_conv -> the pool's context feature) and depth (SE gate
    [-> cat with cost_volumn_net(cost volume) when stereo] -> the HeightNet trunk -> softmax over D) -- and the backward
    of all of it from the pool's depth / context gradients plus the depth loss (lss_heightmap.py:859-897).  The plane-sweep
    cost volume itself is a constant of the step (the reference computes it under no_grad, depthnet.py:405-407);
    cost_volumn_net is trained."""

    def __init__(self, net, device='cuda', loss_weight=0.05, dropout=0.0, seed=0):
        self.stereo = bool(getattr(net, 'stereo', False))
        if self.stereo:
            cv = net.cost_volumn_net
            self.Dcv = cv[0].in_channels
            self.Dcv_pad = (self.Dcv + 63) // 64 * 64
            self.cv1 = _TrainConv(cv[0].weight, cv[0].bias, cv[1], 3, stride=2, cin_pad=self.Dcv_pad, cout_pad=self.Dcv_pad)
            self.cv2 = _TrainConv(cv[2].weight, cv[2].bias, cv[3], 3, stride=2, cin_pad=self.Dcv_pad, cout_pad=self.Dcv_pad)
        super().__init__(net, device, loss_weight, dropout, seed)
        self.context = _TrainConv(net.context_conv.weight, net.context_conv.bias, None, 1,
                                  cout_pad=(net.context_conv.weight.shape[0] + 63) // 64 * 64)
        self.Cctx = net.context_conv.weight.shape[0]
        self.D_bins = self.H_bins

    def _convs(self):
        cs = super()._convs()
        if hasattr(self, 'context'):
            cs.append(self.context)
        if self.stereo:
            cs += [self.cv1, self.cv2]
        return cs

    def _gated(self, x32, gate, name, C_out=None, want32=False):
        N, H, W, C = x32.shape
        out = self._act(name, N, H, W, C_out or C, zero=C_out is not None)
        o32 = self._f32(name + '_32', N, H, W, C) if want32 else None
        _lib.check(_lib.load().dhd_gate_channels(_p(x32), N, H * W, C, _p(gate), _p(out.data), out.ld, out.coff,
                                                 out.part_stride, out.parts, _p(o32), _stream()), 'gate_channels')
        return out, o32

    def forward(self, x, mlp_input, cost_volume=None):
        """x: Act (B*N, C_in, fH, fW); cost_volume: Act (B*N, Dcv_pad, 4fH, 4fW) matching probabilities (stereo) --
        zeros when there is no previous frame.  Returns (softmax depth (B*N, D, fH, fW) NCHW, context (B*N, fH, fW, C))."""
        net = self.net
        N, H, W, C = x.N, x.H, x.W, self.C
        if self.stereo != (cost_volume is not None):
            raise RuntimeError('DepthNet(stereo=%s) called %s a cost volume' % (self.stereo, 'with' if cost_volume is not None else 'without'))
        m_in = mlp_input.reshape(-1, mlp_input.shape[-1]).contiguous().float()
        gd = self._gate_forward(m_in, net.depth_mlp, net.depth_se)
        gc = self._gate_forward(m_in, net.context_mlp, net.context_se)
        nhwc = D.nhwc_strides(C, H, W)
        x32 = self._f32('x32', N, H, W, C)
        self.reduce.forward(x, [dict(act='relu', out_f32=(x32, nhwc))])
        ctx, _ = self._gated(x32, gc['gate'], 'ctx')
        feat = torch.empty(N, H, W, self.Cctx, device=self.device)
        self.context.forward(ctx, [dict(out_f32=(feat, D.nhwc_strides(self.Cctx, H, W)))])
        self.saved = dict(x=x, m_in=m_in, gs=gd, gate=gd['gate'], gc=gc, ctx=ctx)
        if not self.stereo:
            h, h32 = self._gated(x32, gd['gate'], 'h0', want32=True)
            depth = self._trunk_forward(h, h32)
        else:
            cat, _ = self._gated(x32, gd['gate'], 'h0cat', C_out=self.cat_channels)       # channels [0, C) of the zeroed cat buffer
            h2, w2 = (cost_volume.H + 1) // 2, (cost_volume.W + 1) // 2
            cmid = self._act('cv_mid', N, h2, w2, self.Dcv_pad, zero=True)
            self.cv1.forward(cost_volume, [dict(out_act=cmid)])
            self.cv2.forward(cmid, [dict(out_act=cat.slice(C, C + self.Dcv_pad if C + self.Dcv_pad <= self.cat_channels else self.cat_channels))])
            self.saved.update(cv=cost_volume, cmid=cmid)
            depth = self._trunk_forward(cat, None)
        return depth, feat

    def backward(self, depth_grad=None, feat_grad=None, want_dx=True):
        """depth_grad (B*N, D, fH, fW) fp32: gradient at the softmax depth coming from the pool (added to the depth loss
        gradient kept by loss(); either may be absent); feat_grad (B*N, fH, fW, C) fp32: gradient at the context feature."""
        sv, lib = self.saved, _lib.load()
        x = sv['x']
        N, H, W, C = x.N, x.H, x.W, self.C
        HW = H * W
        if depth_grad is not None:
            # through the softmax: p * (g - sum_k p_k g_k), joined with the loss gradient at the logits
            p = sv['height']
            dlog = p * (depth_grad - (p * depth_grad).sum(dim=1, keepdim=True))
            if getattr(self, 'dz', None) is None:
                self.dz = self._act('dz', N, H, W, self.head.cout_pad, zero=True)
                self.dz.data.zero_()
            self.dz.data[..., :self.D_bins].add_(dlog.permute(0, 2, 3, 1))
        dcat = self._trunk_backward()                                    # dL/d (gated depth feature [| cost_volumn_net output])
        self.dz = None
        ws = _workspace(self.device, lib.dhd_sfa_gate_bwd_workspace_bytes(N, HW, C))
        # ---- depth branch gate
        hd = sv['hs'][0]
        dpre = self._act('d_pre', N, H, W, C)
        gsum_d = self._f32('gsum_d', N, C)
        _lib.check(lib.dhd_se_gate_bwd(_p(dcat.data), dcat.ld, dcat.coff, _p(hd.data), hd.ld, hd.coff, C, N, HW,
                                       _p(sv['gate']), _p(dpre.data), dpre.ld, dpre.coff, _p(gsum_d), _p(ws), _stream()),
                   'se_gate_bwd(depth)')
        if self.stereo:
            # ---- cost_volumn_net: weight / bias / BatchNorm gradients only (its input is a constant)
            dcv = dcat.slice(C, C + self.Dcv_pad) if C + self.Dcv_pad <= self.cat_channels else dcat.slice(C, self.cat_channels)
            dmid = self._act('d_cvmid', N, sv['cmid'].H, sv['cmid'].W, self.Dcv_pad)
            _, sums = act_bwd(dcv, None, None, want_sums=True)
            self.cv2.backward(sv['cmid'], dcv, [dict(out_act=dmid)], bias_sums=sums[0])
            _, sums = act_bwd(dmid, None, None, want_sums=True)
            self.cv1.backward(sv['cv'], dmid, None, bias_sums=sums[0])
        # ---- context branch: 1x1 conv + gate
        gsum_c = None
        if feat_grad is not None:
            dfe = self._act('d_feat', N, H, W, self.context.cout_pad, zero=True)
            dfe.data[..., :self.Cctx].copy_(feat_grad.reshape(N, H, W, self.Cctx))
            _, sums = act_bwd(dfe, None, None, want_sums=True)
            dctx = self._act('d_ctx', N, H, W, C)
            self.context.backward(sv['ctx'], dfe, [dict(out_act=dctx)], bias_sums=sums[0])
            dpre_c = self._act('d_pre_c', N, H, W, C)
            gsum_c = self._f32('gsum_c', N, C)
            _lib.check(lib.dhd_se_gate_bwd(_p(dctx.data), dctx.ld, dctx.coff, _p(sv['ctx'].data), sv['ctx'].ld, sv['ctx'].coff,
                                           C, N, HW, _p(sv['gc']['gate']), _p(dpre_c.data), dpre_c.ld, dpre_c.coff,
                                           _p(gsum_c), _p(ws), _stream()), 'se_gate_bwd(context)')
            act_bwd(dpre, None, None, add=dpre_c)                         # both branches meet at the reduce conv's output
        _, sums = act_bwd(dpre, None, None, want_sums=True)
        dx = self._act('d_x', N, H, W, x.C) if want_dx else None
        self.reduce.backward(x, dpre, [dict(out_act=dx)] if want_dx else None, bias_sums=sums[0])
        net = self.net
        self._gate_backward(gsum_d, sv['gs'], sv['m_in'], net.depth_mlp, net.depth_se)
        if gsum_c is not None:
            self._gate_backward(gsum_c, sv['gc'], sv['m_in'], net.context_mlp, net.context_se)
        if self.dropout_p > 0.0:
            self.rng[1:].add_(1)
        return dx


class _TrainConvT:
    """ConvTranspose2d(kernel 2, stride 2) (unet.py:86) for training: forward = four 1x1 GEMMs writing strided
    views; data gradient = ONE stride-2 convolution with the 2x2 taps over the up-sampled gradient; weight gradient =
    the wgrad kernel with the operand roles swapped (the up-sampled gradient is the strided operand)."""
    TAPS = [(0, 0), (0, 1), (1, 0), (1, 1)]

    def __init__(self, m):
        self.m = m
        self.Cin, self.Cout = m.weight.shape[0], m.weight.shape[1]
        self.refresh()

    def refresh(self):
        w = self.m.weight.detach().float()                             # (Cin, Cout, 2, 2)
        self.w_f = [w[:, :, i, j].t().contiguous()[:, None, None, :].to(torch.bfloat16).contiguous()
                    for i in range(2) for j in range(2)]               # (Cout, 1, 1, Cin)
        self.w_b = w.permute(0, 2, 3, 1).reshape(self.Cin, 4, 1, self.Cout).to(torch.bfloat16).contiguous()
        self.bias = self.m.bias.detach() if self.m.bias is not None else None

    def forward(self, x, out):
        ld = out.ld
        for i in range(2):
            for j in range(2):
                D.conv2d(x, self.w_f[2 * i + j], self.Cout, precision='bf16', bias=self.bias,
                         segs=[dict(out_act=out, out_view=(out.H * out.W * ld, 2 * out.W * ld, 2 * ld,
                                                           (i * out.W + j) * ld))])

    def backward(self, x, dy, dx):
        """x: Act (N, Cin, H, W); dy: Act slice on the (>= 2H x 2W) output grid, pad rows / columns already zeroed;
        dx: Act (N, Cin, H, W) written."""
        _, sums = act_bwd(dy, None, None, want_sums=True)
        if self.m.bias is not None:
            _acc(self.m.bias, sums[0][:self.Cout])
        dw = D.conv2d_wgrad(dy, x, self.Cin, taps=self.TAPS, stride=2)           # (Cin, 4, Cout)
        _acc(self.m.weight, dw.view(self.Cin, 2, 2, self.Cout).permute(0, 3, 1, 2))
        D.conv2d(dy, self.w_b, self.Cin, precision='bf16', taps=self.TAPS, stride=2, out_hw=(x.H, x.W),
                 segs=[dict(out_act=dx)])


class _TrainDoubleConv:
    """(conv3x3 -> frozen BN -> ReLU) x 2, unet.py:45-61."""

    def __init__(self, seq):
        self.c1 = _TrainConv(seq[0].weight, None, seq[1], 3)
        self.c2 = _TrainConv(seq[3].weight, None, seq[4], 3)
        self.mid, self.Cout = self.c1.Cout, self.c2.Cout

    def refresh(self):
        self.c1.refresh()
        self.c2.refresh()

    def forward(self, x, out, tmp):
        self.c1.forward(x, [dict(act='relu', out_act=tmp)])
        self.c2.forward(tmp, [dict(act='relu', out_act=out)])

    def backward(self, x, tmp, dy, dtmp, dx_segs):
        """dy: gradient w.r.t. the block output, already masked by its ReLU."""
        self.c2.backward(tmp, dy, [dict(out_act=dtmp)])
        act_bwd(dtmp, tmp, 'relu')
        self.c1.backward(x, dtmp, dx_segs)


class UNetTrainer:
    """UNet (unet.py:6-141, bilinear=False) with frozen BatchNorm: forward with saved activations + backward."""

    def __init__(self, net, device='cuda'):
        self.net, self.device = net, device
        self.inc = _TrainDoubleConv(net.inc.double_conv)
        self.down = [_TrainDoubleConv(getattr(net, 'down%d' % k).maxpool_conv[1].double_conv) for k in range(1, 5)]
        self.upT = [_TrainConvT(getattr(net, 'up%d' % k).up) for k in range(1, 5)]
        self.upC = [_TrainDoubleConv(getattr(net, 'up%d' % k).conv.double_conv) for k in range(1, 5)]
        self.outc = _TrainConv(net.outc.conv.weight, net.outc.conv.bias, None, 1,
                               cout_pad=(net.outc.conv.weight.shape[0] + 63) // 64 * 64)
        self.n_classes = self.outc.Cout
        self._buf = {}

    def refresh(self):
        for m in [self.inc, self.outc] + self.down + self.upT + self.upC:
            m.refresh()

    def _act(self, name, N, H, W, C, zero=False):
        key = (name, N, H, W, C)
        if key not in self._buf:
            a = D.Act.empty(N, H, W, C, 1, self.device)
            if zero:
                a.data.zero_()
            self._buf[key] = a
        return self._buf[key]

    def forward(self, x, out=None):
        N = x.N
        sizes = [(x.H, x.W)]
        for _ in range(4):
            sizes.append((sizes[-1][0] // 2, sizes[-1][1] // 2))
        chans = [self.inc.Cout] + [d.Cout for d in self.down]
        cats = [self._act('cat%d' % k, N, sizes[k][0], sizes[k][1], 2 * chans[k], zero=True) for k in range(4)]
        skip = [cats[k].slice(0, chans[k]) for k in range(4)]
        tmps = {'inc': self._act('tmp0', N, sizes[0][0], sizes[0][1], self.inc.mid)}
        self.inc.forward(x, skip[0], tmps['inc'])
        cur, pooled = skip[0], []
        from .encoders import maxpool2
        for k in range(4):
            H, W = sizes[k + 1]
            pk = self._act('pool%d' % k, N, H, W, chans[k])
            maxpool2(cur, pk)
            pooled.append(pk)
            dst = skip[k + 1] if k < 3 else self._act('bottom', N, H, W, chans[4])
            tmps['d%d' % k] = self._act('tmpd%d' % k, N, H, W, self.down[k].mid)
            self.down[k].forward(pk, dst, tmps['d%d' % k])
            cur = dst
        bottom, decs = cur, []
        for k in range(4):
            lvl = 3 - k
            H, W = sizes[lvl]
            self.upT[k].forward(cur, cats[lvl].slice(chans[lvl], 2 * chans[lvl]))
            dst = self._act('dec%d' % lvl, N, H, W, self.upC[k].Cout)
            tmps['u%d' % k] = self._act('tmpu%d' % lvl, N, H, W, self.upC[k].mid)
            self.upC[k].forward(cats[lvl], dst, tmps['u%d' % k])
            decs.append(dst)
            cur = dst
        if out is None:
            out = self._act('out', N, sizes[0][0], sizes[0][1], self.outc.cout_pad)
        self.outc.forward(cur, [dict(out_act=out)])
        self.saved = dict(x=x, sizes=sizes, chans=chans, cats=cats, skip=skip, tmps=tmps, pooled=pooled, bottom=bottom,
                          decs=decs)
        return out

    def backward(self, dout, dx_f32=None):
        """dout: Act (>= n_classes channels; channels beyond n_classes zero).  dx_f32: optional (tensor, strides)
        receiving dL/dx in fp32 (the layout dhd_mghs_pool_bwd reads); returns the bf16 Act otherwise."""
        sv = self.saved
        x, sizes, chans, cats, skip, tmps = sv['x'], sv['sizes'], sv['chans'], sv['cats'], sv['skip'], sv['tmps']
        pooled, bottom, decs = sv['pooled'], sv['bottom'], sv['decs']
        N = x.N
        lib = _lib.load()
        _, sums = act_bwd(dout, None, None, want_sums=True)
        y4 = decs[3]
        d = self._act('g_dec0', N, y4.H, y4.W, y4.C)
        self.outc.backward(y4, dout, [dict(out_act=d)], bias_sums=sums[0])
        dskip = [None] * 4
        for k in range(3, -1, -1):                                   # up4 .. up1
            lvl = 3 - k
            H, W = sizes[lvl]
            blk, y = self.upC[k], decs[k]
            act_bwd(d, y, 'relu')
            dcat = self._act('g_cat%d' % lvl, N, H, W, 2 * chans[lvl])
            blk.backward(cats[lvl], tmps['u%d' % k], d, self._act('g_tmpu%d' % lvl, N, H, W, blk.mid),
                         [dict(out_act=dcat)])
            dskip[lvl] = dcat.slice(0, chans[lvl])
            src = decs[k - 1] if k > 0 else bottom                    # the map the transposed conv up-sampled
            dup = dcat.slice(chans[lvl], 2 * chans[lvl])
            if H != 2 * src.H:                                        # pad row / column of the odd-sized level
                dcat.data[:, 2 * src.H:, :, chans[lvl]:] = 0
            if W != 2 * src.W:
                dcat.data[:, :, 2 * src.W:, chans[lvl]:] = 0
            d = self._act('g_src%d' % lvl, N, src.H, src.W, src.C)
            self.upT[k].backward(src, dup, d)
        # d = dL/d bottom; the contracting path, deepest first
        for k in range(3, -1, -1):
            blk = self.down[k]
            y = bottom if k == 3 else skip[k + 1]
            H, W = y.H, y.W
            if k < 3:                                                  # decoder skip gradient + deeper level's
                act_bwd(d, y, 'relu', add=dskip[k + 1])
            else:
                act_bwd(d, y, 'relu')
            dpool = self._act('g_pool%d' % k, N, H, W, chans[k])
            blk.backward(pooled[k], tmps['d%d' % k], d, self._act('g_tmpd%d' % k, N, H, W, blk.mid), [dict(out_act=dpool)])
            xin = skip[k]
            d = self._act('g_x%d' % k, N, xin.H, xin.W, chans[k])
            _lib.check(lib.dhd_maxpool2_bwd(_p(xin.data), xin.ld, xin.coff, _p(dpool.data), dpool.ld, dpool.coff, N,
                                            xin.H, xin.W, chans[k], _p(d.data), d.ld, d.coff, _stream()), 'maxpool2_bwd')
        act_bwd(d, skip[0], 'relu', add=dskip[0])
        if dx_f32 is not None:
            segs = [dict(out_f32=dx_f32)]
            dx = None
        else:
            dx = self._act('g_in', N, x.H, x.W, x.C)
            segs = [dict(out_act=dx)]
        self.inc.backward(x, tmps['inc'], d, self._act('g_tmp0', N, x.H, x.W, self.inc.mid), segs)
        return dx


class CustomResNetTrainer:
    """CustomResNet (resnet.py:10-80, block_type='Basic') with frozen BatchNorm: forward + backward."""

    def __init__(self, net, device='cuda'):
        self.device = device
        self.stages = []
        for stage in net.layers:
            blocks = []
            for b in stage:
                st = b.conv1.stride[0]
                c1 = _TrainConv(b.conv1.weight, None, b.bn1, 3, stride=st)
                c2 = _TrainConv(b.conv2.weight, None, b.bn2, 3)
                ds = _TrainConv(b.downsample.weight, b.downsample.bias, None, 3, stride=st) if b.downsample is not None else None
                blocks.append((c1, c2, ds))
            self.stages.append(blocks)
        self.output_ids = list(net.backbone_output_ids)
        self._buf = {}

    def refresh(self):
        for blocks in self.stages:
            for c1, c2, ds in blocks:
                c1.refresh()
                c2.refresh()
                if ds is not None:
                    ds.refresh()

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

    def forward(self, x):
        self.saved, feats, x32 = [], [], None
        for si, blocks in enumerate(self.stages):
            for bi, (c1, c2, ds) in enumerate(blocks):
                tag = '%d_%d' % (si, bi)
                oH, oW = (x.H, x.W) if c1.stride == 1 else ((x.H + 1) // 2, (x.W + 1) // 2)
                nh = D.nhwc_strides(c2.Cout, oH, oW)
                t = self._act('t' + tag, x.N, oH, oW, c1.Cout)
                c1.forward(x, [dict(act='relu', out_act=t)])
                if ds is not None:
                    idn = self._f32('idn' + tag, x.N, oH, oW, c2.Cout)
                    ds.forward(x, [dict(out_f32=(idn, nh))])
                else:
                    idn = x32
                out, out32 = self._act('o' + tag, x.N, oH, oW, c2.Cout), self._f32('o32' + tag, x.N, oH, oW, c2.Cout)
                c2.forward(t, [dict(act='relu', out_act=out, out_f32=(out32, nh))], residual=(idn, nh[:3]))
                self.saved.append((x, t, out))
                x, x32 = out, out32
            if si in self.output_ids:
                feats.append(x)
        return feats

    def backward(self, dfeats, dx_f32=None):
        """dfeats: {stage index: Act gradient of that stage's output} (bf16, not yet masked)."""
        idx, d = len(self.saved), None
        for si in range(len(self.stages) - 1, -1, -1):
            for bi in range(len(self.stages[si]) - 1, -1, -1):
                idx -= 1
                c1, c2, ds = self.stages[si][bi]
                x, t, out = self.saved[idx]
                tag = '%d_%d' % (si, bi)
                last = bi == len(self.stages[si]) - 1
                extra = dfeats.get(si) if last else None
                if d is None:
                    d = extra
                    act_bwd(d, out, 'relu')
                else:
                    act_bwd(d, out, 'relu', add=extra)
                # d = gradient at the block's pre-ReLU sum: goes to conv2's branch and to the identity / downsample
                dt = self._act('g_t' + tag, t.N, t.H, t.W, t.C)
                c2.backward(t, d, [dict(out_act=dt)])
                act_bwd(dt, t, 'relu')
                dxin = self._act('g_x' + tag, x.N, x.H, x.W, x.C)
                first = idx == 0
                if first and dx_f32 is not None and ds is not None:
                    # the network input: both branches into one fp32 tensor through the residual input
                    raise NotImplementedError('fp32 input gradient of a stride-2 first block: use the bf16 Act')
                c1.backward(x, dt, [dict(out_act=dxin)])
                if ds is not None:
                    _, sums = act_bwd(d, None, None, want_sums=True)
                    dxid = self._act('g_id' + tag, x.N, x.H, x.W, x.C)
                    ds.backward(x, d, [dict(out_act=dxid)], bias_sums=sums[0])
                    act_bwd(dxin, None, None, add=dxid)            # sum of the two paths (no mask: next block / caller masks)
                else:
                    act_bwd(dxin, None, None, add=d)               # identity path
                d = dxin
        return d


class FPNLSSTrainer:
    """FPN_LSS (lss_fpn.py:11-74, lateral=None, extra_upsample=2) with frozen BatchNorm: forward + backward."""

    def __init__(self, neck, device='cuda'):
        self.device = device
        self.idx = tuple(neck.input_feature_index)
        self.scale, self.scale2 = int(neck.up.scale_factor), int(neck.up2[0].scale_factor)
        self.c1 = _TrainConv(neck.conv[0].weight, None, neck.conv[1], 3)
        self.c2 = _TrainConv(neck.conv[3].weight, None, neck.conv[4], 3)
        self.c3 = _TrainConv(neck.up2[1].weight, None, neck.up2[2], 3)
        self.c4 = _TrainConv(neck.up2[4].weight, neck.up2[4].bias, None, 1)
        self._buf = {}

    def refresh(self):
        for c in (self.c1, self.c2, self.c3, self.c4):
            c.refresh()

    _act = CustomResNetTrainer._act
    _f32 = CustomResNetTrainer._f32

    def forward(self, feats, out=None):
        from .encoders import upsample_bilinear
        x2, x1 = feats[self.idx[0]], feats[self.idx[1]]
        N, H, W = x2.N, x2.H, x2.W
        cat = self._act('cat', N, H, W, x2.C + x1.C)
        upsample_bilinear(x2, cat.slice(0, x2.C))
        upsample_bilinear(x1, cat.slice(x2.C, x2.C + x1.C))
        t = self._act('t', N, H, W, self.c1.Cout)
        self.c1.forward(cat, [dict(act='relu', out_act=t)])
        u = self._act('u', N, H, W, self.c2.Cout)
        self.c2.forward(t, [dict(act='relu', out_act=u)])
        H2, W2 = H * self.scale2, W * self.scale2
        v = self._act('v', N, H2, W2, u.C)
        upsample_bilinear(u, v)
        w = self._act('w', N, H2, W2, self.c3.Cout)
        self.c3.forward(v, [dict(act='relu', out_act=w)])
        if out is None:
            out = self._act('out', N, H2, W2, self.c4.Cout)
        self.c4.forward(w, [dict(out_act=out)])
        self.saved = (x2, x1, cat, t, u, v, w)
        return out

    def backward(self, dout):
        """Returns {feature index: Act gradient} for the two inputs."""
        x2, x1, cat, t, u, v, w = self.saved
        lib = _lib.load()
        N = x2.N
        _, sums = act_bwd(dout, None, None, want_sums=True)
        dw_ = self._act('g_w', N, w.H, w.W, w.C)
        self.c4.backward(w, dout, [dict(out_act=dw_)], bias_sums=sums[0])
        act_bwd(dw_, w, 'relu')
        dv = self._act('g_v', N, v.H, v.W, v.C)
        self.c3.backward(v, dw_, [dict(out_act=dv)])
        du = self._act('g_u', N, u.H, u.W, u.C)
        self._upsample_bwd(dv, u, v.H, v.W, du, 'g_u32')
        act_bwd(du, u, 'relu')
        dt = self._act('g_t', N, t.H, t.W, t.C)
        self.c2.backward(t, du, [dict(out_act=dt)])
        act_bwd(dt, t, 'relu')
        dcat = self._act('g_cat', N, cat.H, cat.W, cat.C)
        self.c1.backward(cat, dt, [dict(out_act=dcat)])
        d2 = dcat.slice(0, x2.C)                                    # identity "up-sampling": the slice is the gradient
        up = dcat.slice(x2.C, x2.C + x1.C)
        d1 = self._act('g_x1', N, x1.H, x1.W, x1.C)
        self._upsample_bwd(up, x1, cat.H, cat.W, d1, 'g_x1_32')
        return {self.idx[0]: d2, self.idx[1]: d1}

    def _upsample_bwd(self, dy, src, out_H, out_W, dx, scratch):
        """dx (bf16 Act, the shape of `src`) = backward of the bilinear up-sampling of `src` to out_H x out_W at dy."""
        lib = _lib.load()
        N = src.N
        if UPSAMPLE_BWD_GATHER:
            _lib.check(lib.dhd_upsample_bilinear_bwd_gather(_p(dy.data), dy.ld, dy.coff, N, src.H, src.W, src.C, out_H, out_W,
                                                            _p(dx.data), dx.ld, dx.coff, None, _stream()), 'upsample_bwd_gather')
            return
        d32 = self._f32(scratch, N, src.H, src.W, src.C)
        _lib.check(lib.dhd_upsample_bilinear_bwd(_p(dy.data), dy.ld, dy.coff, N, src.H, src.W, src.C, out_H, out_W, _p(d32),
                                                 _stream()), 'upsample_bwd')
        _lib.check(lib.dhd_add_rowvec(_p(d32), None, N, src.H * src.W, src.C, _p(dx.data), dx.ld, dx.coff, _stream()), 'f32->bf16')
