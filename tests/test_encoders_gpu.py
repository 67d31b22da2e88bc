"""GPU parity of the BEV / voxel encoder path: the new convolution modes (stride 2 through a strided TMA box,
ConvTranspose2d(2, 2) as four 1x1 GEMMs writing a strided view), MaxPool2d(2) and bilinear up-sampling against
torch, and the UNet / CustomResNet / FPN_LSS engines against outputs of the unmodified reference classes
(tests/golden/encoders.npz).  fp32 mode (6-term split-bf16): atol 2e-4 of the tensor's max."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def close(got, want, tol=2e-4):
    scale = float(want.abs().max())
    err = float((got.cpu() - want).abs().max())
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize('N,Cin,Cout,H,W', [(2, 64, 128, 40, 56), (1, 128, 256, 25, 25), (3, 256, 64, 13, 7)])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_conv_stride2(cuda_lib, N, Cin, Cout, H, W, precision):
    from dhd_b200 import dense as D
    g = torch.Generator().manual_seed(Cin + H)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05
    b = torch.randn(Cout, generator=g)
    if precision == 'bf16':
        x, w = x.bfloat16().float(), w.bfloat16().float()
    want = F.conv2d(x, w, b, stride=2, padding=1)
    parts = D.PRECISIONS[precision][0]
    oH, oW = (H + 1) // 2, (W + 1) // 2
    out = torch.empty(N, oH, oW, Cout, device='cuda')
    D.conv2d(D.pack_input(x.cuda(), parts), D.pack_weight(w.cuda(), parts), Cout, ksize=3, precision=precision,
             bias=b.cuda(), stride=2, segs=[dict(out_f32=(out, D.nhwc_strides(Cout, oH, oW)))])
    close(out.permute(0, 3, 1, 2), want, 2e-4 if precision == 'fp32' else 2e-3)


@pytest.mark.parametrize('H,W,pad', [(12, 12, 1), (10, 14, 0)])
def test_conv_transpose_2x2_strided_view(cuda_lib, H, W, pad):
    from dhd_b200 import dense as D
    from dhd_b200.encoders import _ConvT2x2
    torch.manual_seed(H)
    m = torch.nn.ConvTranspose2d(128, 64, kernel_size=2, stride=2)
    x = torch.randn(2, 128, H, W)
    want = F.pad(m(x).detach(), [0, pad, 0, pad])
    cat = D.Act.empty(2, 2 * H + pad, 2 * W + pad, 128, 3, 'cuda')       # [other 64 | up 64]
    cat.data.zero_()
    _ConvT2x2(m, 'fp32', 'cuda')(D.pack_input(x.cuda(), 3), cat.slice(64, 128))
    close(cat.slice(64, 128).float(), want)
    assert float(cat.slice(0, 64).float().abs().max()) == 0.0               # the neighbouring slice is untouched


def test_maxpool_and_bilinear_upsample(cuda_lib):
    from dhd_b200 import dense as D
    from dhd_b200.encoders import maxpool2, upsample_bilinear
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 64, 25, 50, generator=g)
    xa = D.pack_input(x.cuda(), 3)
    out = D.Act.empty(2, 12, 25, 64, 3, 'cuda')
    close(maxpool2(xa, out).float(), F.max_pool2d(x, 2), 1e-6)
    for s in (2, 4):
        up = D.Act.empty(2, 25 * s, 50 * s, 128, 3, 'cuda')
        up.data.zero_()
        upsample_bilinear(xa, up.slice(64, 128))
        close(up.slice(64, 128).float(), F.interpolate(x, scale_factor=s, mode='bilinear', align_corners=True), 1e-5)


def _gold():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'encoders.npz'))


def _modules():
    import projects.mmdet3d_plugin  # noqa: F401
    from oracle import dense_oracle as DO
    from oracle import make_golden_encoders as ME
    from projects.mmdet3d_plugin.models.backbones import CustomResNet, UNet
    from projects.mmdet3d_plugin.models.necks import FPN_LSS
    mods = dict(unet=UNet(256, 64), resnet=CustomResNet(64, num_channels=[128, 256, 512]), fpn=FPN_LSS(640, 256))
    for k, m in mods.items():
        m.load_state_dict(DO.seeded_state_dict(m, ME.SEEDS[k]))
        m.eval().cuda()
    return ME, mods


def test_unet_matches_reference_fixture(cuda_lib):
    ME, mods = _modules()
    xu, _ = ME.inputs()
    close(mods['unet'](xu.cuda()), torch.from_numpy(_gold()['unet']))


def test_bev_encoder_matches_reference_fixture(cuda_lib):
    ME, mods = _modules()
    _, xb = ME.inputs()
    gold = _gold()
    feats = mods['resnet'](xb.cuda(), return_act=True)
    for i, f in enumerate(feats):
        close(f.float(), torch.from_numpy(gold['feat%d' % i]))
    close(mods['fpn'](feats), torch.from_numpy(gold['fpn']))
