"""Host side of the dense layers of the hot path (ctypes over the C-ABI; torch only owns the
device memory and the stream).

Activations travel between layers as ``Act``: NHWC bf16 with up to three "parts" stacked on
the channel axis (x = x0 + x1 + x2, each part bf16) -- 1 part is plain bf16, 3 parts carry a
full fp32 value.  ``conv2d`` runs one implicit-GEMM convolution on the tcgen05 tensor cores
(csrc/conv_igemm.cu) with the BatchNorm / bias / residual / activation / gate epilogue fused.

Precision modes (``PRECISIONS``): the number of bf16 x bf16 MMAs issued per product term
  'bf16'   1 part,  1 term   -- speed mode
  'bf16x3' 2 parts, 3 terms  -- ~2^-16 relative
  'fp32'   3 parts, 6 terms  -- fp32-grade; the mode the 1e-4 logits parity is checked in
"""
import ctypes

import torch

from . import _lib

ACT = {None: 0, 'none': 0, 'relu': 1, 'sigmoid': 2, 'softplus': 3, 'softmax': 4}
PRECISIONS = {
    'bf16': (1, [(0, 0)]),
    'bf16x3': (2, [(0, 0), (0, 1), (1, 0)]),
    'fp32': (3, [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)]),
}
MAX_TAPS, MAX_TERMS, MAX_SEGS = 9, 6, 2


class ConvSeg(ctypes.Structure):
    _fields_ = [
        ('c_lo', ctypes.c_int32), ('c_hi', ctypes.c_int32), ('act', ctypes.c_int32),
        ('out_f32', ctypes.c_void_p),
        ('f32_sN', ctypes.c_int64), ('f32_sY', ctypes.c_int64), ('f32_sX', ctypes.c_int64),
        ('f32_sC', ctypes.c_int64),
        ('out_b16', ctypes.c_void_p),
        ('b16_ld', ctypes.c_int32), ('b16_coff', ctypes.c_int32), ('b16_parts', ctypes.c_int32),
        ('b16_part_stride', ctypes.c_int32),
        ('b16_sN', ctypes.c_int64), ('b16_sY', ctypes.c_int64), ('b16_sX', ctypes.c_int64),
    ]


class ConvDesc(ctypes.Structure):
    _fields_ = [
        ('N', ctypes.c_int32), ('H', ctypes.c_int32), ('W', ctypes.c_int32),
        ('Cin', ctypes.c_int32), ('Cout', ctypes.c_int32), ('taps', ctypes.c_int32),
        ('tap_dy', ctypes.c_int32 * MAX_TAPS), ('tap_dx', ctypes.c_int32 * MAX_TAPS),
        ('bw', ctypes.c_int32), ('bh', ctypes.c_int32),
        ('in_', ctypes.c_void_p),
        ('in_ld', ctypes.c_int32), ('in_coff', ctypes.c_int32), ('in_part_stride', ctypes.c_int32),
        ('weight', ctypes.c_void_p), ('w_parts', ctypes.c_int32), ('n_terms', ctypes.c_int32),
        ('term_a', ctypes.c_int32 * MAX_TERMS), ('term_b', ctypes.c_int32 * MAX_TERMS),
        ('scale', ctypes.c_void_p), ('bias', ctypes.c_void_p), ('img_bias', ctypes.c_void_p),
        ('img_gate', ctypes.c_void_p), ('residual', ctypes.c_void_p),
        ('res_sN', ctypes.c_int64), ('res_sY', ctypes.c_int64), ('res_sX', ctypes.c_int64),
        ('n_seg', ctypes.c_int32), ('seg', ConvSeg * MAX_SEGS),
        ('stride', ctypes.c_int32), ('in_H', ctypes.c_int32), ('in_W', ctypes.c_int32),
        ('mix_x', ctypes.c_void_p), ('mix_ld', ctypes.c_int32), ('mix_coff', ctypes.c_int32),
        ('mix_parts', ctypes.c_int32), ('mix_part_stride', ctypes.c_int32), ('mix_a1', ctypes.c_void_p),
        ('w_image_rows', ctypes.c_int32), ('res_b16_ld', ctypes.c_int32), ('res_b16_coff', ctypes.c_int32),
        ('res_b16', ctypes.c_void_p), ('stat_partial', ctypes.c_void_p),
    ]


class PredictorTailDesc(ctypes.Structure):
    """struct dhd_predictor_tail_desc (include/dhd_b200.h)."""
    _fields_ = [
        ('N', ctypes.c_int32), ('H', ctypes.c_int32), ('W', ctypes.c_int32),
        ('K1', ctypes.c_int32), ('N1', ctypes.c_int32), ('Dz', ctypes.c_int32), ('n_cls', ctypes.c_int32),
        ('in_ld', ctypes.c_int32), ('in_coff', ctypes.c_int32), ('transpose_xy', ctypes.c_int32),
        ('in_', ctypes.c_void_p), ('w1', ctypes.c_void_p), ('b1', ctypes.c_void_p),
        ('w2', ctypes.c_void_p), ('b2', ctypes.c_void_p), ('logits', ctypes.c_void_p), ('occ', ctypes.c_void_p),
        ('hidden', ctypes.c_void_p), ('hidden_ld', ctypes.c_int32), ('hidden_coff', ctypes.c_int32),
    ]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return t.data_ptr() if t is not None else None


class Act:
    """NHWC bf16 activation with `parts` split-bf16 parts: tensor (N, H, W, parts*C)."""

    def __init__(self, data, C, parts):
        self.data, self.C, self.parts = data, C, parts
        self.N, self.H, self.W = data.shape[:3]
        self.ld = data.shape[3]
        self.coff = 0
        self.part_stride = C
        self.mean = None          # optional (N, C) fp32 per-image channel means made by the producer

    @staticmethod
    def empty(N, H, W, C, parts, device):
        return Act(torch.empty(N, H, W, parts * C, dtype=torch.bfloat16, device=device), C, parts)

    def slice(self, c_lo, c_hi):
        """A channel sub-range view (same storage), e.g. one branch of a concatenation."""
        a = Act.__new__(Act)
        a.data, a.C, a.parts = self.data, c_hi - c_lo, self.parts
        a.N, a.H, a.W, a.ld = self.N, self.H, self.W, self.ld
        a.coff, a.part_stride = self.coff + c_lo, self.part_stride
        a.mean = None
        return a

    def float(self):
        """fp32 NCHW value (sum of the parts) -- for tests / hand-off to torch code."""
        x = self.data.float().view(self.N, self.H, self.W, -1)
        v = 0
        for p in range(self.parts):
            lo = self.coff + p * self.part_stride
            v = v + x[..., lo:lo + self.C]
        return v.permute(0, 3, 1, 2).contiguous()


def split_bf16(x, parts):
    """fp32 tensor -> list of `parts` bf16 tensors whose sum reproduces x to ~2^(-8*parts)."""
    out, r = [], x.float()
    for _ in range(parts):
        h = r.to(torch.bfloat16)
        out.append(h)
        r = r - h.float()
    return out


def pack_weight(w, parts):
    """(Cout, Cin, kh, kw) or (Cout, Cin) fp32 -> bf16 [Cout][taps][parts][Cin] contiguous
    (done once when a module is built, not on the hot path)."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    Cout, Cin, kh, kw = w.shape
    w = w.detach().float().permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin)       # [Cout][tap][Cin]
    ps = split_bf16(w, parts)
    return torch.stack(ps, dim=2).contiguous()                                     # [Cout][tap][part][Cin]


def pack_input(x, parts):
    """(N, C, H, W) fp32 -> Act (C padded to a multiple of 64 with zeros).  Uses the library's
    pack kernel when present (csrc/layout.cu), torch otherwise is NOT a fallback: raises."""
    lib = _lib.load()
    N, C, H, W = x.shape
    Cp = (C + 63) // 64 * 64
    x = x.contiguous().float()
    out = Act.empty(N, H, W, Cp, parts, x.device)
    if Cp != C:
        out.data.zero_()
    _lib.check(lib.dhd_pack_nchw_to_nhwc(ctypes.c_void_p(x.data_ptr()), N, C, H, W,
                                         ctypes.c_void_p(out.data.data_ptr()), out.ld, 0, Cp, parts,
                                         _stream()), 'pack_nchw_to_nhwc')
    return out


_MEAN_WS = {}


def mean_workspace(device, N, C, HW):
    """Scratch of the deterministic channel-mean reductions (block partials), one buffer per (device, stream)."""
    need = _lib.load().dhd_mean_workspace_bytes(N, C, HW)
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _MEAN_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _MEAN_WS[key] = ws
    return ws


def pack_any(x, parts, want_mean=False):
    """Logical (N, C, H, W) fp32 tensor in either memory format -> Act."""
    if x.dim() != 4:
        raise ValueError('expected a (N, C, H, W) tensor')
    if not x.is_contiguous() and x.permute(0, 2, 3, 1).is_contiguous():
        return pack_nhwc(x.permute(0, 2, 3, 1), parts, want_mean=want_mean)
    return pack_input(x, parts)


def pack_nhwc(x, parts, want_mean=False):
    """(N, H, W, C) contiguous fp32 -> Act.  want_mean=True also produces the per-image channel
    means in the same pass (attached as `act.mean`, consumed by the SFA squeeze)."""
    lib = _lib.load()
    N, H, W, C = x.shape
    Cp = (C + 63) // 64 * 64
    x = x.contiguous().float()
    out = Act.empty(N, H, W, Cp, parts, x.device)
    if Cp != C:
        out.data.zero_()
    if want_mean and C % 8 == 0 and 256 % (C // 8) == 0:
        mean = torch.empty(N, C, device=x.device)
        ws = mean_workspace(x.device, N, C, H * W)
        _lib.check(lib.dhd_split_nhwc_mean(ctypes.c_void_p(x.data_ptr()), N, H * W, C,
                                           ctypes.c_void_p(out.data.data_ptr()), out.ld, 0, Cp, parts,
                                           ctypes.c_void_p(mean.data_ptr()), ctypes.c_void_p(ws.data_ptr()),
                                           _stream()), 'split_nhwc_mean')
        out.mean = mean
        return out
    _lib.check(lib.dhd_split_nhwc(ctypes.c_void_p(x.data_ptr()), N * H * W, C,
                                  ctypes.c_void_p(out.data.data_ptr()), out.ld, 0, Cp, parts,
                                  _stream()), 'split_nhwc')
    return out


def tile_box(H, W):
    """(bw, bh) with bw*bh == 128 minimising the padded area of an H x W image."""
    best = None
    for bw in (4, 8, 16, 32, 64, 128):
        bh = 128 // bw
        area = ((W + bw - 1) // bw * bw) * ((H + bh - 1) // bh * bh)
        if best is None or area < best[0]:
            best = (area, bw, bh)
    return best[1], best[2]


def conv2d(x, weight, Cout, ksize=1, dilation=1, precision='fp32', scale=None, bias=None,
           img_bias=None, img_gate=None, residual=None, segs=None, stride=1, sfa_mix=None, taps=None,
           out_hw=None, residual_act=None, image_weights=False, defer=False, stats=None):
    """One fused convolution.  x: Act; weight: pack_weight() result with PRECISIONS[precision]
    parts; segs: list of dicts {c_lo, c_hi, act, out_f32 (tensor, strides (sN,sY,sX,sC)),
    out_act (Act or Act.slice), out_view (sN, sY, sX, offset): pixel strides / start offset in bf16
    elements when out_act is written as a strided view (ConvTranspose2d phases)}.  stride=2: 3x3 / pad 1
    down-sampling convolution, output ceil(H/2) x ceil(W/2).  Outputs are written in place.
    residual_act: single-part bf16 Act added before the activation (bf16 speed mode's identity path);
    image_weights=True: `weight` is [N*Cout][taps][parts][Cin], image n convolves with its own Cout rows.
    defer=True: nothing is launched; returns (descriptor, tensors to keep alive) for conv2d_batch()."""
    parts, terms = PRECISIONS[precision]
    if x.parts < parts or weight.shape[2] != parts:
        raise ValueError('activation has %d parts, weight %d, precision %s needs %d' %
                         (x.parts, weight.shape[2], precision, parts))
    d = ConvDesc()
    oH, oW = (x.H, x.W) if stride == 1 else ((x.H + 1) // 2, (x.W + 1) // 2)
    if out_hw is not None:                   # an output grid smaller than the input allows (top-left part)
        oH, oW = out_hw
    d.N, d.H, d.W = x.N, oH, oW
    d.stride = stride
    if stride != 1 or (oH, oW) != (x.H, x.W):
        d.in_H, d.in_W = x.H, x.W
    d.Cin, d.Cout = x.C, Cout
    tap_list = taps                          # explicit input offsets [(dy, dx), ...] (must contain (0, 0))
    taps = ksize * ksize if tap_list is None else len(tap_list)
    if image_weights:
        if weight.shape[0] != Cout * x.N:
            raise ValueError('per-image weights need N*Cout = %d rows, got %d' % (Cout * x.N, weight.shape[0]))
        d.w_image_rows = Cout
    if weight.shape[0] != (Cout * x.N if image_weights else Cout) or weight.shape[1] != taps or weight.shape[3] != x.C:
        raise ValueError('weight shape %s does not match Cout=%d taps=%d Cin=%d' %
                         (tuple(weight.shape), Cout, taps, x.C))
    d.taps = taps
    r = ksize // 2
    for t in range(taps):
        if tap_list is not None:
            d.tap_dy[t], d.tap_dx[t] = tap_list[t]
        else:
            d.tap_dy[t] = (t // ksize - r) * dilation
            d.tap_dx[t] = (t % ksize - r) * dilation
    d.bw, d.bh = tile_box(oH, oW)
    d.in_ = x.data.data_ptr()
    d.in_ld, d.in_coff, d.in_part_stride = x.ld, x.coff, x.part_stride
    d.weight = weight.data_ptr()
    d.w_parts = parts
    d.n_terms = len(terms)
    for i, (a, b) in enumerate(terms):
        d.term_a[i], d.term_b[i] = a, b
    keep = [x.data, weight]
    for name, t in (('scale', scale), ('bias', bias), ('img_bias', img_bias), ('img_gate', img_gate)):
        if t is not None:
            if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                raise ValueError(name + ' must be a contiguous fp32 CUDA tensor')
            setattr(d, name, t.data_ptr())
            keep.append(t)
    if residual is not None:
        rt, (sN, sY, sX) = residual
        d.residual = rt.data_ptr()
        d.res_sN, d.res_sY, d.res_sX = sN, sY, sX
        keep.append(rt)
    if residual_act is not None:
        ra = residual_act
        if (ra.N, ra.H, ra.W) != (x.N, oH, oW) or ra.C < Cout or residual is not None:
            raise ValueError('residual_act does not match the layer (or an fp32 residual is also given)')
        d.res_b16, d.res_b16_ld, d.res_b16_coff = ra.data.data_ptr(), ra.ld, ra.coff
        keep.append(ra.data)
    if sfa_mix is not None:                  # (x Act [bev | vox], a1 (N, Cout) fp32): SFA blend in the epilogue
        mx, a1 = sfa_mix
        if (mx.N, mx.H, mx.W) != (x.N, oH, oW) or mx.C < 2 * Cout or a1.shape != (x.N, Cout) or a1.dtype != torch.float32:
            raise ValueError('sfa_mix operands do not match the layer')
        d.mix_x, d.mix_ld, d.mix_coff = mx.data.data_ptr(), mx.ld, mx.coff
        d.mix_parts, d.mix_part_stride = mx.parts, mx.part_stride
        d.mix_a1 = a1.data_ptr()
        keep += [mx.data, a1]
    d.n_seg = len(segs)
    for i, s in enumerate(segs):
        g = d.seg[i]
        g.c_lo, g.c_hi = s.get('c_lo', 0), s.get('c_hi', Cout)
        g.act = ACT[s.get('act')]
        if s.get('out_f32') is not None:
            t, st = s['out_f32']
            g.out_f32 = t.data_ptr()
            g.f32_sN, g.f32_sY, g.f32_sX, g.f32_sC = st
            keep.append(t)
        if s.get('out_act') is not None:
            a = s['out_act']
            view = s.get('out_view')
            if a.C < g.c_hi - g.c_lo or (view is None and (a.N, a.H, a.W) != (x.N, oH, oW)):
                raise ValueError('output activation does not match the layer')
            g.out_b16 = a.data.data_ptr() + (0 if view is None else 2 * view[3])
            g.b16_ld, g.b16_coff, g.b16_parts, g.b16_part_stride = a.ld, a.coff, a.parts, a.part_stride
            if view is not None:
                g.b16_sN, g.b16_sY, g.b16_sX = view[0], view[1], view[2]
            keep.append(a.data)
    if stats is not None:                    # (2, Cout) fp32: [sum, sum of squares] of the bf16 output per channel
        lib = _lib.load()
        rows = lib.dhd_conv2d_stat_rows(ctypes.byref(d))
        part = _stat_workspace(x.data.device, rows * 2 * Cout * 4)
        d.stat_partial = part.data_ptr()
        keep.append(part)
        if defer:
            raise ValueError('stats with defer is not supported')
        _lib.check(lib.dhd_conv2d_fwd(ctypes.byref(d), _stream()), 'conv2d_fwd')
        if isinstance(stats, str):               # 'partial': the caller finishes (e.g. fused with the BatchNorm formulas)
            return part, rows
        _lib.check(lib.dhd_colsum_finish(part.data_ptr(), rows, 2 * Cout, stats.data_ptr(), _stream()), 'colsum_finish')
        return keep
    if defer:
        return d, keep
    _lib.check(_lib.load().dhd_conv2d_fwd(ctypes.byref(d), _stream()), 'conv2d_fwd')
    return keep


_STAT_WS = {}


def _stat_workspace(device, nbytes):
    key = (str(device), torch.cuda.current_stream().cuda_stream)
    ws = _STAT_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 22), dtype=torch.uint8, device=device)
        _STAT_WS[key] = ws
    return ws


MAX_BATCH = 4


def conv2d_batch(deferred):
    """Launch the convolutions prepared with conv2d(..., defer=True) MAX_BATCH at a time in one persistent kernel each
    (dhd_conv2d_fwd_batch): for independent layers that are individually too small to fill the GPU."""
    deferred = [x for x in deferred if x is not None]
    lib = _lib.load()
    for i in range(0, len(deferred), MAX_BATCH):
        chunk = deferred[i:i + MAX_BATCH]
        arr = (ConvDesc * len(chunk))()
        for j, (d, _) in enumerate(chunk):
            ctypes.memmove(ctypes.addressof(arr[j]), ctypes.addressof(d), ctypes.sizeof(ConvDesc))
        _lib.check(lib.dhd_conv2d_fwd_batch(arr, len(chunk), _stream()), 'conv2d_fwd_batch')
    return [k for _, k in deferred]


def predictor_tail(x, w1, b1, w2, b2, Dz, n_cls, logits=None, occ=None, transpose_xy=True, hidden=None):
    """Fused Linear + Softplus + Linear (+ per-z argmax) of the occupancy head (dhd_predictor_tail, bf16 operands).
    x: Act (B, K1, H, W), part 0 is read; w1 / w2: pack_weight(.., 1) results; b1 / b2 fp32.
    logits: fp32 (B, W, H, Dz*n_cls) and / or occ: uint8 (B, W, H, Dz), written in place."""
    d = PredictorTailDesc()
    d.N, d.H, d.W = x.N, x.H, x.W
    d.K1, d.N1, d.Dz, d.n_cls = x.C, w1.shape[0], Dz, n_cls
    if w1.shape[3] != x.C or w2.shape[3] != w1.shape[0] or w2.shape[0] != Dz * n_cls or w1.shape[1:3] != (1, 1) or \
            w2.shape[1:3] != (1, 1) or w1.dtype != torch.bfloat16 or w2.dtype != torch.bfloat16:
        raise ValueError('predictor_tail: weights must be bf16 [N1][1][1][K1] and [Dz*n_cls][1][1][N1]')
    d.in_ld, d.in_coff, d.transpose_xy = x.ld, x.coff, int(bool(transpose_xy))
    d.in_, d.w1, d.w2 = x.data.data_ptr(), w1.data_ptr(), w2.data_ptr()
    for name, t, n in (('b1', b1, w1.shape[0]), ('b2', b2, w2.shape[0])):
        if t is not None:
            if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
                raise ValueError(name + ' must be a contiguous fp32 vector of %d elements' % n)
            setattr(d, name, t.data_ptr())
    npix = x.N * x.H * x.W
    if logits is not None:
        if logits.dtype != torch.float32 or not logits.is_contiguous() or logits.numel() != npix * Dz * n_cls:
            raise ValueError('logits must be contiguous fp32 with B*H*W*Dz*n_cls elements')
        d.logits = logits.data_ptr()
    if occ is not None:
        if occ.dtype != torch.uint8 or not occ.is_contiguous() or occ.numel() != npix * Dz:
            raise ValueError('occ must be contiguous uint8 with B*H*W*Dz elements')
        d.occ = occ.data_ptr()
    if hidden is not None:                   # Act (N, H, W, >= N1): the Softplus output, saved for the backward
        if (hidden.N, hidden.H, hidden.W) != (x.N, x.H, x.W) or hidden.C < w1.shape[0]:
            raise ValueError('hidden activation does not match the layer')
        d.hidden, d.hidden_ld, d.hidden_coff = hidden.data.data_ptr(), hidden.ld, hidden.coff
    _lib.check(_lib.load().dhd_predictor_tail(ctypes.byref(d), _stream()), 'predictor_tail')


def nchw_strides(C, H, W):
    return (C * H * W, W, 1, H * W)


def nhwc_strides(C, H, W):
    return (H * W * C, W * C, C, 1)


# ------------------------------------------------------------------------------ backward
class WgradDesc(ctypes.Structure):
    """struct dhd_wgrad_desc (include/dhd_b200.h)."""
    _fields_ = [
        ('N', ctypes.c_int32), ('H', ctypes.c_int32), ('W', ctypes.c_int32),
        ('Cin', ctypes.c_int32), ('Cout', ctypes.c_int32), ('taps', ctypes.c_int32),
        ('tap_dy', ctypes.c_int32 * MAX_TAPS), ('tap_dx', ctypes.c_int32 * MAX_TAPS),
        ('bw', ctypes.c_int32), ('bh', ctypes.c_int32),
        ('x', ctypes.c_void_p), ('x_ld', ctypes.c_int32), ('x_coff', ctypes.c_int32),
        ('dy', ctypes.c_void_p), ('dy_ld', ctypes.c_int32), ('dy_coff', ctypes.c_int32),
        ('scale', ctypes.c_void_p), ('dw', ctypes.c_void_p), ('partial', ctypes.c_void_p),
        ('accumulate', ctypes.c_int32), ('x_stride', ctypes.c_int32), ('x_H', ctypes.c_int32), ('x_W', ctypes.c_int32),
        ('dw_torch', ctypes.c_int32), ('dw_cin_total', ctypes.c_int32), ('dw_cin_lo', ctypes.c_int32),
        ('dw_cin_used', ctypes.c_int32),
    ]


_WGRAD_WS = {}


def conv2d_wgrad(x, dy, Cout, ksize=1, dilation=1, scale=None, out=None, accumulate=False, stride=1, taps=None,
                 grad=None):
    """Weight gradient of a stride-1 'same' convolution: x, dy are Acts (part 0 is used: bf16
    operands, fp32 accumulation).  Returns dw fp32 (Cout, ksize*ksize, Cin) -- the tap-major layout
    of pack_weight(); `weight_grad_to_torch` turns it into (Cout, Cin, kh, kw).
    grad=(tensor, cin_lo, cin_used): ADD the gradient straight into `tensor`, the contiguous fp32 gradient of an
    nn.Conv2d / nn.Linear weight (Cout, Cin_total[, kh, kw]): input channel ci < cin_used goes to column cin_lo + ci."""
    if x.N != dy.N or dy.C < Cout or (stride == 1 and (x.H, x.W) != (dy.H, dy.W)):
        raise ValueError('x / dy shapes do not match')
    d = WgradDesc()
    d.N, d.H, d.W, d.Cin, d.Cout = x.N, dy.H, dy.W, x.C, Cout          # the pixel loop runs over dy's grid
    if stride != 1:                          # x sampled at (stride*y + tap, stride*x + tap) of its own grid
        d.x_stride, d.x_H, d.x_W = stride, x.H, x.W
    tap_list = taps
    taps = ksize * ksize if tap_list is None else len(tap_list)
    d.taps = taps
    r = ksize // 2
    for t in range(taps):
        if tap_list is not None:
            d.tap_dy[t], d.tap_dx[t] = tap_list[t]
        else:
            d.tap_dy[t] = (t // ksize - r) * dilation
            d.tap_dx[t] = (t % ksize - r) * dilation
    d.bw, d.bh = tile_box(dy.H, dy.W)
    d.x, d.x_ld, d.x_coff = x.data.data_ptr(), x.ld, x.coff
    d.dy, d.dy_ld, d.dy_coff = dy.data.data_ptr(), dy.ld, dy.coff
    if scale is not None:
        d.scale = scale.data_ptr()
    if grad is not None:
        out, cin_lo, cin_used = grad
        if out.dtype != torch.float32 or not out.is_contiguous() or out.shape[0] != Cout or \
                out.numel() != Cout * out.shape[1] * taps:
            raise ValueError('grad must be the contiguous fp32 (Cout, Cin_total[, kh, kw]) gradient tensor')
        d.dw_torch, d.dw_cin_total, d.dw_cin_lo, d.dw_cin_used = 1, out.shape[1], cin_lo, cin_used
        accumulate = True
    elif out is None:
        out = torch.empty(Cout, taps, x.C, device=x.data.device)
        accumulate = False
    d.dw = out.data_ptr()
    d.accumulate = int(accumulate)
    lib = _lib.load()
    need = lib.dhd_conv2d_wgrad_workspace_bytes(ctypes.byref(d))
    key = (x.data.device, torch.cuda.current_stream().cuda_stream)
    ws = _WGRAD_WS.get(key)
    if ws is None or ws.numel() < need:         # one scratch buffer per (device, stream), grown on demand
        ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=x.data.device)
        _WGRAD_WS[key] = ws
    d.partial = ws.data_ptr()
    _lib.check(lib.dhd_conv2d_wgrad(ctypes.byref(d), _stream()), 'conv2d_wgrad')
    return out


def weight_grad_to_torch(dw, ksize):
    """(Cout, taps, Cin) -> (Cout, Cin, kh, kw), the shape of nn.Conv2d.weight.grad."""
    Cout, taps, Cin = dw.shape
    return dw.view(Cout, ksize, ksize, Cin).permute(0, 3, 1, 2).contiguous()


def pack_weight_dgrad(w, parts=1):
    """Forward weight (Cout, Cin, kh, kw) -> the packed weight of the data-gradient convolution:
    dx = conv(dy, w'), w'[ci][tap'][co] = w[co][tap][ci] with the taps mirrored."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    return pack_weight(w.detach().transpose(0, 1).flip(2, 3), parts)
