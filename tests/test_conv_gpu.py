"""GPU parity tests of the tcgen05 implicit-GEMM convolution (csrc/conv_igemm.cu) and the
streaming layout kernels (csrc/layout.cu) against torch fp32 references of the same ops.

Tolerances (written here, per precision mode): 'fp32' = 6-term split-bf16, max abs error
<= 2e-5 * (|x| . |w|) row-norm bound -- i.e. fp32-accumulation grade; 'bf16x3' <= 1e-4 of the
bound; 'bf16' <= 2^-7 of the bound (inputs rounded to bf16)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {'fp32': 2e-5, 'bf16x3': 2e-4, 'bf16': 1.6e-2}


def _bound(x, w, dilation, pad):
    """sum_k |x_k| |w_k| per output element: the scale rounding errors are relative to."""
    return F.conv2d(x.abs(), w.abs(), padding=pad, dilation=dilation)


def _run(x, w, ksize, dilation, precision, **kw):
    from dhd_b200 import dense as D
    parts, _ = D.PRECISIONS[precision]
    xa = D.pack_input(x, parts)
    wp = D.pack_weight(w, parts).cuda()
    N, C, H, W = x.shape
    Cout = w.shape[0]
    out = torch.full((N, Cout, H, W), float('nan'), device='cuda')
    D.conv2d(xa, wp, Cout, ksize=ksize, dilation=dilation, precision=precision,
             segs=[dict(act=kw.pop('act', None), out_f32=(out, D.nchw_strides(Cout, H, W)))], **kw)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3', 'bf16'])
@pytest.mark.parametrize('shape', [
    # N, Cin, Cout, H, W, ksize, dilation
    (2, 64, 128, 16, 44, 1, 1),
    (3, 256, 256, 16, 44, 3, 1),
    (2, 128, 65, 16, 44, 3, 6),
    (1, 256, 256, 16, 44, 3, 18),
    (1, 64, 108, 8, 16, 1, 1),
    (1, 128, 288, 40, 24, 3, 1),
])
def test_conv_matches_torch(cuda_lib, shape, precision):
    N, Cin, Cout, H, W, k, dil = shape
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    x = torch.randn(N, Cin, H, W, generator=g).cuda()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).cuda()
    pad = dil * (k // 2)
    ref = F.conv2d(x.double(), w.double(), padding=pad, dilation=dil).float()
    out = _run(x, w, k, dil, precision)
    assert torch.isfinite(out).all(), 'unwritten / non-finite outputs'
    err = ((out - ref).abs() / _bound(x, w, dil, pad).clamp_min(1e-6)).max().item()
    assert err <= TOL[precision], 'relative-to-bound error %.3g > %.3g' % (err, TOL[precision])


def test_pack_unpack_roundtrip(cuda_lib):
    from dhd_b200 import dense as D, _lib
    import ctypes
    x = torch.randn(2, 100, 9, 13, device='cuda')
    a = D.pack_input(x, 3)
    assert a.C == 128 and a.data.shape == (2, 9, 13, 3 * 128)
    back = a.float()[:, :100]
    assert (back - x).abs().max().item() <= 1e-6 * x.abs().max().item()
    assert a.float()[:, 100:].abs().max().item() == 0
    out = torch.empty(2, 100, 9, 13, device='cuda')
    _lib.check(_lib.load().dhd_unpack_nhwc_to_nchw(
        ctypes.c_void_p(a.data.data_ptr()), a.ld, 0, a.part_stride, 3, 2, 100, 9, 13,
        ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
        'unpack')
    assert torch.equal(out, back)


def test_conv_epilogue_bn_residual_relu_gate(cuda_lib):
    """scale/bias (folded BN) + per-image bias + residual + ReLU + per-image gate, bf16 split out."""
    from dhd_b200 import dense as D
    N, Cin, Cout, H, W = 2, 64, 192, 16, 44
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, Cin, H, W, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / 24).cuda()
    scale = (torch.rand(Cout, generator=g) + 0.5).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    ib = torch.randn(N, Cout, generator=g).cuda()
    gate = torch.rand(N, Cout, generator=g).cuda()
    res = torch.randn(N, H, W, Cout, generator=g).cuda()
    ref = F.conv2d(x.double(), w.double(), padding=1).float() * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    ref = ref + ib.view(N, Cout, 1, 1) + res.permute(0, 3, 1, 2)
    ref = torch.relu(ref) * gate.view(N, Cout, 1, 1)
    xa = D.pack_input(x, 3)
    wp = D.pack_weight(w, 3).cuda()
    oa = D.Act.empty(N, H, W, Cout, 3, 'cuda')
    of = torch.empty(N, H, W, Cout, device='cuda')
    D.conv2d(xa, wp, Cout, ksize=3, precision='fp32', scale=scale, bias=bias, img_bias=ib,
             img_gate=gate, residual=(res, (H * W * Cout, W * Cout, Cout)),
             segs=[dict(act='relu', out_act=oa, out_f32=(of, D.nhwc_strides(Cout, H, W)))])
    torch.cuda.synchronize()
    assert torch.allclose(of.permute(0, 3, 1, 2), ref, rtol=1e-4, atol=1e-4)
    assert torch.allclose(oa.float(), of.permute(0, 3, 1, 2), rtol=0, atol=1e-6)


def test_conv_two_segments_softmax(cuda_lib):
    """depth_net-style head: channels [0,44) softmaxed into an NCHW plane, [44,108) raw NHWC."""
    from dhd_b200 import dense as D
    N, Cin, H, W, Dd, C = 3, 256, 16, 44, 44, 64
    g = torch.Generator().manual_seed(4)
    x = torch.randn(N, Cin, H, W, generator=g).cuda()
    w = (torch.randn(Dd + C, Cin, 1, 1, generator=g) / 8).cuda()
    b = torch.randn(Dd + C, generator=g).cuda()
    ref = F.conv2d(x.double(), w.double(), b.double()).float()   # (cuDNN fp32 convs default to TF32)
    depth_ref = ref[:, :Dd].softmax(dim=1)
    feat_ref = ref[:, Dd:].permute(0, 2, 3, 1)
    xa = D.pack_input(x, 3)
    wp = D.pack_weight(w, 3).cuda()
    depth = torch.empty(N, Dd, H, W, device='cuda')
    feat = torch.empty(N, H, W, C, device='cuda')
    D.conv2d(xa, wp, Dd + C, precision='fp32', bias=b, segs=[
        dict(c_lo=0, c_hi=Dd, act='softmax', out_f32=(depth, D.nchw_strides(Dd, H, W))),
        dict(c_lo=Dd, c_hi=Dd + C, out_f32=(feat, D.nhwc_strides(C, H, W)))])
    torch.cuda.synchronize()
    assert torch.allclose(depth, depth_ref, rtol=1e-4, atol=1e-6)
    assert torch.allclose(feat, feat_ref, rtol=1e-4, atol=1e-4)
