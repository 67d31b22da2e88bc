"""GPU parity of the convolution backward kernels (C-ABI dhd_conv2d_wgrad, and dhd_conv2d_fwd on the
mirrored weight for the data gradient) against torch autograd of F.conv2d on the same bf16-rounded
operands (fp32 accumulation on both sides: differences are summation order only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _act(x_nchw):
    from dhd_b200 import dense as D
    return D.pack_input(x_nchw, 1)


CASES = [
    # N, H, W, Cin, Cout, k, dil
    (2, 16, 44, 256, 256, 3, 1),     # HeightNet trunk
    (2, 16, 44, 256, 256, 3, 6),     # ASPP branch (dilated: dead taps at the borders)
    (1, 40, 40, 512, 256, 1, 1),     # SFA shortcut
    (1, 40, 40, 256, 512, 1, 1),     # predicter[0]
    (1, 40, 40, 512, 288, 1, 1),     # predicter[2]: Cout not a multiple of 128
    (3, 13, 21, 64, 128, 3, 1),      # ragged: partial boxes on both image edges
    (2, 16, 44, 256, 64, 1, 1),      # Cout < 128
]


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,dil', CASES)
def test_wgrad_matches_torch(cuda_lib, N, H, W, Cin, Cout, k, dil):
    from dhd_b200 import dense as D
    g = torch.Generator().manual_seed(N * 1000 + Cin + Cout + k + dil)
    x = torch.randn(N, Cin, H, W, generator=g).bfloat16().float().cuda()
    dy = (torch.randn(N, Cout, H, W, generator=g) * 0.1).bfloat16().float().cuda()
    w = torch.zeros(Cout, Cin, k, k, device='cuda', requires_grad=True)
    y = torch.nn.functional.conv2d(x, w, padding=dil * (k // 2), dilation=dil)
    (want,) = torch.autograd.grad(y, w, dy)
    got = D.weight_grad_to_torch(D.conv2d_wgrad(_act(x), _act(dy), Cout, ksize=k, dilation=dil), k)
    scale = float(want.abs().max())
    assert torch.allclose(got, want, rtol=2e-3, atol=2e-4 * scale), float((got - want).abs().max()) / scale
    # scale / accumulate epilogue of the reduction
    s = torch.rand(Cout, device='cuda') + 0.5
    acc = want.clone().view(Cout, Cin, k * k).permute(0, 2, 1).contiguous()
    D.conv2d_wgrad(_act(x), _act(dy), Cout, ksize=k, dilation=dil, scale=s, out=acc, accumulate=True)
    want2 = want + want * s.view(-1, 1, 1, 1)
    assert torch.allclose(D.weight_grad_to_torch(acc, k), want2, rtol=2e-3, atol=4e-4 * scale)


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,dil', CASES[:4] + CASES[5:6])
def test_dgrad_matches_torch(cuda_lib, N, H, W, Cin, Cout, k, dil):
    from dhd_b200 import dense as D
    g = torch.Generator().manual_seed(7 + Cin + Cout + k + dil)
    w = (torch.randn(Cout, Cin, k, k, generator=g) * 0.05).bfloat16().float().cuda()
    dy = torch.randn(N, Cout, H, W, generator=g).bfloat16().float().cuda()
    x = torch.zeros(N, Cin, H, W, device='cuda', requires_grad=True)
    y = torch.nn.functional.conv2d(x, w, padding=dil * (k // 2), dilation=dil)
    (want,) = torch.autograd.grad(y, x, dy)
    out = torch.empty(N, H, W, Cin, device='cuda')
    D.conv2d(_act(dy), D.pack_weight_dgrad(w), Cin, ksize=k, dilation=dil, precision='bf16',
             segs=[dict(out_f32=(out, D.nhwc_strides(Cin, H, W)))])
    got = out.permute(0, 3, 1, 2)
    scale = float(want.abs().max())
    assert torch.allclose(got, want, rtol=2e-3, atol=2e-4 * scale), float((got - want).abs().max()) / scale


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(2, 37, 53, 128, 256, 3), (4, 16, 44, 256, 96, 1), (1, 200, 200, 64, 320, 3)])
def test_conv_epilogue_batchnorm_statistics(cuda_lib, shape):
    """dhd_conv_desc.stat_partial + dhd_colsum_finish: per-channel sum / sum of squares of the layer's bf16 output
    (BatchNorm batch statistics) == the same sums taken from the stored output, in both the one-CTA and the CTA-pair
    kernel; partial tiles at the image border and a channel count that is not a multiple of 64 included."""
    from dhd_b200 import _lib
    from dhd_b200 import dense as D
    N, H, W, Cin, Cout, k = shape
    g = torch.Generator().manual_seed(3)
    xa = D.pack_input(torch.randn(N, Cin, H, W, generator=g).cuda(), 1)
    wq = D.pack_weight((torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).cuda(), 1)
    bias = torch.randn(Cout, generator=g).cuda()
    lib = _lib.load()
    prev = lib.dhd_conv_pair_mode(-1)
    try:
        for mode in (0, 2):
            lib.dhd_conv_pair_mode(mode)
            out = D.Act.empty(N, H, W, (Cout + 63) // 64 * 64, 1, 'cuda')
            out.data.zero_()
            sums = torch.full((2, Cout), float('nan'), device='cuda')
            D.conv2d(xa, wq, Cout, ksize=k, precision='bf16', bias=bias, segs=[dict(out_act=out)], stats=sums)
            torch.cuda.synchronize()
            y = out.data[..., :Cout].double()
            want = torch.stack([y.sum((0, 1, 2)), (y * y).sum((0, 1, 2))])
            err = (sums.double() - want).abs() / (want.abs() + 1.0)
            assert float(err.max()) < 2e-5, (mode, float(err.max()))
    finally:
        lib.dhd_conv_pair_mode(prev)
