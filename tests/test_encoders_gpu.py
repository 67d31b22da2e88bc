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


def _dhd_s_model_cfg():
    """The `model` dict of projects/configs/DHD/DHD-S.py:41-155 without the image backbone / FPN."""
    c = 64
    grid = {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [-1, 5.4, 6.4], 'depth': [1.0, 45.0, 1.0]}
    mg = lambda z: {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': z, 'depth': [1.0, 45.0, 0.5]}
    return dict(
        type='DHD',
        img_view_transformer=dict(type='MGHS', grid_config=grid, input_size=(256, 704),
                                  height_range=[round(-1.0 + 0.1 * i, 1) for i in range(65)], height_interval=0.1,
                                  mask_1_grid=mg([-1, 0.6, 0.4]), mask_2_grid=mg([0.6, 2.2, 0.4]), mask_3_grid=mg([2.2, 5.4, 0.4]),
                                  mask_range=[-1.0, 0.6, 2.2, 5.4], loss_height_weight=0.1, in_channels=256, out_channels=c,
                                  sid=False, collapse_z=True, downsample=16),
        img_bev_encoder_backbone=dict(type='CustomResNet', numC_input=c, num_channels=[c * 2, c * 4, c * 8]),
        img_bev_encoder_neck=dict(type='FPN_LSS', in_channels=c * 8 + c * 2, out_channels=256),
        img_voxel_encoder0_backbone=dict(type='UNet', n_channels=c * 4, n_classes=64), img_voxel_encoder0_neck=dict(type='Identity'),
        img_voxel_encoder1_backbone=dict(type='UNet', n_channels=c * 4, n_classes=128), img_voxel_encoder1_neck=dict(type='Identity'),
        img_voxel_encoder2_backbone=dict(type='UNet', n_channels=c * 8, n_classes=64), img_voxel_encoder2_neck=dict(type='Identity'),
        mix=dict(type='SFA', in_channels=512, out_channels=256),
        occ_head=dict(type='predictor', in_dim=256, out_dim=256, Dz=16, use_mask=True, num_classes=18, use_predicter=True,
                      class_balance=True, loss_occ=dict(type='CrossEntropyLoss', use_sigmoid=False, ignore_index=255, loss_weight=1.0)))


def test_dhd_detector_from_config_image_features_to_occupancy(cuda_lib):
    """The DHD detector built from the DHD-S model config through the registry (view transformer, BEV encoder,
    three voxel encoders, SFA, occupancy head): image features -> occupancy logits at the config's full grid,
    against the oracle's restatement of every stage after the view transform fed with the detector's own pooled
    BEV tensors (fp32 mode)."""
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200 import compat as C
    from dhd_b200 import synth
    from oracle import dense_oracle as DO
    model = C.DETECTORS.build(_dhd_s_model_cfg()).eval()
    model.load_state_dict(DO.seeded_state_dict(model, 77))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    B, N = 1, 6
    rig = synth.synthetic_rig(B, N, (256, 704), seed=5)
    x = DO.seeded_tensor((B, N, 256, 16, 44), 78)
    cams = [t.cuda() for t in rig]
    occ, depth, height = model.forward_hot_path(x.cuda(), cams)
    assert occ.shape == (B, 200, 200, 16, 18)
    bev, _, _, lo, mid, hi = model.view_transform(x.cuda(), cams)
    sub = lambda p: {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    with torch.no_grad():
        x2d = DO.fpn_lss_forward(sub('img_bev_encoder_neck.'), DO.custom_resnet_forward(sub('img_bev_encoder_backbone.'), bev.cpu().contiguous()))
        x3d = [DO.unet_forward(sub('img_voxel_encoder%d.' % i), t.cpu().contiguous()) for i, t in enumerate((lo, mid, hi))]
        want = DO.predictor_forward(sub('occ_head.'), DO.sfa_forward(sub('mix.'), torch.cat([x2d] + x3d, dim=1)))
    close(occ, want, 5e-4)


def test_pipeline_encoders_on_parallel_streams_equal_sequential(cuda_lib):
    """HotPathStep(encoders=True) runs the three UNets on side streams: same bits as one stream."""
    from dhd_b200 import synth
    from dhd_b200.pipeline import HotPathStep
    cfg, B = synth.DHD_S, 1
    step = HotPathStep(cfg, B, precision='bf16', use_graph=False, encoders=True)
    host = step.make_host_inputs(synth.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=3), seed=3)
    step.alloc_static(host)
    step.upload(host)
    step._front()
    step._pool()
    a = step._encode(parallel=True).data.clone()
    torch.cuda.synchronize()
    b = step._encode(parallel=False).data.clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b) and float(a.float().abs().max()) > 0
    step._back()
    torch.cuda.synchronize()
    assert int(step.occ.max()) <= 17
