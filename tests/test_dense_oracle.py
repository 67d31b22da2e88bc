"""CPU tests of the dense-path oracle and of the plugin boundary:
(a) oracle/dense_oracle.py reproduces the fixture made from the REAL reference modules;
(b) where /root/reference exists, it also matches those modules directly;
(c) the plugin's modules expose the reference's parameter names and shapes;
(d) the reference's DHD-*.py configs load unchanged and build the hot-path modules."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import dense_oracle as DO
from oracle import make_golden_dense as MG
from oracle import ref_loader

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'dense_modules.npz')


def plugin_modules():
    import projects.mmdet3d_plugin  # noqa: F401  (registers)
    from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
    from projects.mmdet3d_plugin.models.model_utils.depthnet import HeightNet
    from projects.mmdet3d_plugin.models.necks.mix import SFA
    hn = HeightNet(256, 256, 65).eval()
    sfa = SFA(512, 256).eval()
    head = predictor(in_dim=256, out_dim=256, Dz=16, num_classes=18, use_predicter=True,
                     class_balance=False, loss_occ=None).eval()
    return hn, sfa, head, torch.nn.Conv2d(256, 108, 1)


def seeded(mods):
    return [DO.seeded_state_dict(m, s) for m, s in zip(mods, (21, 22, 23, 24))]


def test_seeded_weights_reproduce_the_fixture_hashes():
    gold = np.load(GOLD)
    sds = seeded(plugin_modules())
    for name, sd in zip(('heightnet', 'sfa', 'predictor', 'depth_net'), sds):
        assert MG.sha_sd(sd) == str(gold['sha_' + name]), name
    x, mlp, bev = MG.inputs()
    assert hashlib.sha256(b''.join(t.numpy().tobytes() for t in (x, mlp, bev))).hexdigest() == str(gold['input_sha'])


def test_oracle_matches_reference_fixture():
    gold = np.load(GOLD)
    hn_sd, sfa_sd, head_sd, dn_sd = seeded(plugin_modules())
    x, mlp, bev = MG.inputs()
    with torch.no_grad():
        height = DO.heightnet_forward(hn_sd, x, mlp)
        y = torch.nn.functional.conv2d(x, dn_sd['weight'], dn_sd['bias'])
        fused = DO.sfa_forward(sfa_sd, bev)
        occ = DO.predictor_forward(head_sd, fused)
    for got, key in ((height, 'height'), (y, 'depth_net'), (fused, 'sfa'), (occ, 'occ')):
        ref = torch.from_numpy(gold[key])
        assert got.shape == ref.shape
        assert torch.allclose(got, ref, rtol=1e-5, atol=2e-5), \
            '%s: max abs diff %.3g' % (key, (got - ref).abs().max())


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_oracle_matches_real_reference_modules():
    hn, sfa, head, dn = MG.build_reference_modules()
    x, mlp, bev = MG.inputs()
    for m, s in zip((hn, sfa, head, dn), (21, 22, 23, 24)):
        m.load_state_dict(DO.seeded_state_dict(m, s))
    with torch.no_grad():
        assert torch.allclose(DO.heightnet_forward(hn.state_dict(), x, mlp), hn(x, mlp), rtol=1e-5, atol=2e-5)
        fused = sfa(bev)
        assert torch.allclose(DO.sfa_forward(sfa.state_dict(), bev), fused, rtol=1e-5, atol=2e-5)
        assert torch.allclose(DO.predictor_forward(head.state_dict(), fused), head(fused), rtol=1e-5, atol=2e-5)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_plugin_state_dict_names_match_reference():
    ref = MG.build_reference_modules()
    ours = plugin_modules()
    for r, o in zip(ref, ours):
        rs, os_ = r.state_dict(), o.state_dict()
        assert list(rs.keys()) == list(os_.keys())
        for k in rs:
            assert rs[k].shape == os_[k].shape, k
    ns = ref_loader.load_reference()
    from dhd_b200.compat import Config
    cfg = Config.fromfile(os.path.join(ref_loader.load_reference_configs(), 'DHD-S.py'))
    vt = dict(cfg.model.img_view_transformer)
    vt.pop('type')
    import projects.mmdet3d_plugin.models.necks.lss_heightmap as LH
    ours_vt, ref_vt = LH.MGHS(**vt), ns.MGHS(**vt)
    assert list(ours_vt.state_dict().keys()) == list(ref_vt.state_dict().keys())
    assert torch.equal(ours_vt.frustum, ref_vt.frustum)
    for a in ('grid_lower_bound', 'grid_interval', 'grid_size'):
        assert torch.equal(getattr(ours_vt, a), getattr(ref_vt, a))
    assert ours_vt.D == ref_vt.D and ours_vt.H == ref_vt.H
    g = torch.Generator().manual_seed(0)
    args = [torch.randn(2, 6, 4, 4, generator=g), torch.randn(2, 6, 4, 4, generator=g), torch.randn(2, 6, 3, 3, generator=g),
            torch.randn(2, 6, 3, 3, generator=g), torch.randn(2, 6, 3, generator=g), torch.randn(2, 3, 3, generator=g)]
    assert torch.equal(ours_vt.get_mlp_input(*args), ref_vt.get_mlp_input(*args))
    # height-loss path (GT min-pool, binning, BCE) against the reference's own code
    gt_d = torch.where(torch.rand(1, 6, 256, 704, generator=g) < 0.02, 1 + 50 * torch.rand(1, 6, 256, 704, generator=g), torch.zeros(1))
    gt_h = torch.where(gt_d > 0, -1.2 + 7 * torch.rand(1, 6, 256, 704, generator=g), torch.zeros(1))
    h = torch.rand(6, 65, 16, 44, generator=g).softmax(1)
    for m in (ours_vt, ref_vt):       # the reference computes the loss after view_transform left mask_3_grid behind
        m.grid_config = m.mask_3_grid
    assert torch.equal(ours_vt.get_downsampled_gt_depth(gt_d), ref_vt.get_downsampled_gt_depth(gt_d))
    assert torch.equal(ours_vt.get_downsampled_gt_height(gt_h), ref_vt.get_downsampled_gt_height(gt_h))
    assert torch.allclose(ours_vt.get_height_loss(gt_d, gt_h, h), ref_vt.get_height_loss(gt_d, gt_h, h), rtol=1e-6)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
@pytest.mark.parametrize('name', ['DHD-S.py', 'DHD-M.py', 'DHD-L.py'])
def test_reference_configs_load_unchanged(name):
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200 import compat as C
    cfg = C.Config.fromfile(os.path.join(ref_loader.load_reference_configs(), name))
    assert cfg.plugin and cfg.plugin_dir == 'projects/mmdet3d_plugin/'
    assert cfg.dist_params.backend == 'nccl'            # from the un-vendored _base_ fallback
    assert cfg.model.type in ('DHD', 'DHD_stereo')
    if name == 'DHD-S.py':
        model = C.build_model(cfg.model)
        vt = model.img_view_transformer
        assert (vt.D, vt.H, vt.out_channels) == (44, 65, 64)
        assert type(model.mix).__name__ == 'SFA' and type(model.occ_head).__name__ == 'predictor'
        assert model.occ_head.cls_weights.shape == (18,)
    cfg.merge_from_dict({'model.img_view_transformer.accelerate': True})
    assert cfg.model.img_view_transformer.accelerate is True


def test_dense_modules_refuse_cpu_tensors():
    hn, sfa, head, _ = plugin_modules()
    with pytest.raises(RuntimeError, match='CUDA'):
        sfa(torch.zeros(1, 512, 8, 16))
    with pytest.raises(RuntimeError, match='CUDA'):
        head(torch.zeros(1, 256, 8, 16))
    with pytest.raises(RuntimeError, match='CUDA'):
        hn(torch.zeros(1, 256, 8, 16), torch.zeros(1, 1, 27))


# ---------------------------------------------------------------- BEV / voxel encoders
def _encoder_state_dicts():
    import projects.mmdet3d_plugin  # noqa: F401
    from oracle import make_golden_encoders as ME
    from projects.mmdet3d_plugin.models.backbones import CustomResNet, UNet
    from projects.mmdet3d_plugin.models.necks import FPN_LSS
    mods = dict(unet=UNet(256, 64), resnet=CustomResNet(64, num_channels=[128, 256, 512]), fpn=FPN_LSS(640, 256))
    return ME, {k: DO.seeded_state_dict(m, ME.SEEDS[k]) for k, m in mods.items()}, mods


def test_encoder_oracle_matches_reference_fixture():
    """oracle restatement of UNet / CustomResNet / FPN_LSS against outputs of the unmodified reference classes
    (tests/golden/encoders.npz, made by oracle/make_golden_encoders.py)."""
    ME, sds, _ = _encoder_state_dicts()
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'encoders.npz'))
    for k, sd in sds.items():
        assert MG.sha_sd(sd) == str(gold['sha_' + k]), 'seeded weights differ from the fixture'
    xu, xb = ME.inputs()
    with torch.no_grad():
        yu = DO.unet_forward(sds['unet'], xu)
        feats = DO.custom_resnet_forward(sds['resnet'], xb)
        yf = DO.fpn_lss_forward(sds['fpn'], feats)
    for got, key in ((yu, 'unet'), (feats[0], 'feat0'), (feats[1], 'feat1'), (feats[2], 'feat2'), (yf, 'fpn')):
        want = torch.from_numpy(gold[key])
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max())), key


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_encoder_oracle_and_plugin_names_match_real_reference():
    ME, sds, mods = _encoder_state_dicts()
    refs = dict(zip(('unet', 'resnet', 'fpn'), ME.build_reference_modules()))
    for k in refs:
        assert list(refs[k].state_dict().keys()) == list(mods[k].state_dict().keys()), k
        assert all(refs[k].state_dict()[n].shape == mods[k].state_dict()[n].shape for n in refs[k].state_dict()), k
        refs[k].load_state_dict(sds[k])
    xu, xb = ME.inputs()
    with torch.no_grad():
        assert torch.equal(refs['unet'](xu), DO.unet_forward(sds['unet'], xu))
        fr = refs['resnet'](xb)
        fo = DO.custom_resnet_forward(sds['resnet'], xb)
        assert all(torch.equal(a, b) for a, b in zip(fr, fo))
        assert torch.equal(refs['fpn'](fr), DO.fpn_lss_forward(sds['fpn'], fo))


def test_depthnet_oracle_matches_reference_fixture():
    """oracle.dense_oracle.depthnet_forward / stereo_sampling_grid / stereo_cost_volume against the fixture made by
    the UNMODIFIED reference DepthNet (oracle/make_golden_depthnet.py): DHD-M form, DHD-L stereo form, first frame."""
    from oracle import make_golden_depthnet as MGD
    gold = np.load(os.path.join(os.path.dirname(GOLD), 'depthnet.npz'))
    x, mlp, prev, curr = MGD.inputs()
    assert hashlib.sha256(b''.join(t.numpy().tobytes() for t in (x, mlp, prev, curr))).hexdigest() == str(gold['input_sha'])
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.model_utils.depthnet import DepthNet
    D = MGD.n_depth()
    mono = DepthNet(MGD.C_IN, MGD.C_MID_MONO, MGD.C_CTX, D, use_dcn=True, use_aspp=True)
    st = DepthNet(MGD.C_IN, MGD.C_IN, MGD.C_CTX, D, use_dcn=False, aspp_mid_channels=32, stereo=True, bias=MGD.BIAS)
    sdm, sds = DO.seeded_state_dict(mono, 41), DO.seeded_state_dict(st, 42)     # plugin parameter names == reference's
    assert MG.sha_sd(sdm) == str(gold['sha_mono']) and MG.sha_sd(sds) == str(gold['sha_stereo'])
    m = MGD.stereo_metas(prev, curr)
    H, W = MGD.INPUT
    with torch.no_grad():
        grid = DO.stereo_sampling_grid(m['frustum'], m['k2s_sensor'], m['intrins'], m['post_rots'], m['post_trans'], H, W)
        assert np.abs(grid.numpy() - gold['grid']).max() <= 1e-6
        cv = DO.stereo_cost_volume(prev, curr, grid, D, MGD.BIAS)
        assert np.abs(cv.numpy() - gold['cost_volume']).max() <= 1e-6
        for got, key in ((DO.depthnet_forward(sdm, x, mlp), 'mono'), (DO.depthnet_forward(sds, x, mlp, cv), 'stereo'),
                         (DO.depthnet_forward(sds, x, mlp, torch.zeros_like(cv)), 'stereo_first')):
            ref = torch.from_numpy(gold[key])
            assert torch.allclose(got, ref, rtol=1e-5, atol=2e-5), '%s: max abs diff %.3g' % (key, (got - ref).abs().max())


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_depthnet_fixture_is_what_the_reference_produces_now():
    """Re-run the unmodified reference DepthNet (stereo) in this container and compare with the committed fixture."""
    from oracle import make_golden_depthnet as MGD
    gold = np.load(os.path.join(os.path.dirname(GOLD), 'depthnet.npz'))
    ns = ref_loader.load_reference()
    st = MGD.build(ns, True)
    st.load_state_dict(DO.seeded_state_dict(st, 42))
    x, mlp, prev, curr = MGD.inputs()
    with torch.no_grad():
        y = st(x, mlp, MGD.stereo_metas(prev, curr))
    assert torch.allclose(y, torch.from_numpy(gold['stereo']), rtol=1e-5, atol=2e-5)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_dhdl_view_transformer_kwargs_are_the_reference_config():
    """dhd_b200.synth.DHD_L_VIEW_TRANSFORMER (used on the GPU box, where the reference tree is absent) is exactly
    model.img_view_transformer of the unchanged projects/configs/DHD/DHD-L.py."""
    from dhd_b200 import compat as C
    from dhd_b200 import synth
    cfg = C.Config.fromfile(os.path.join(ref_loader.load_reference_configs(), 'DHD-L.py'))
    ref = cfg.model.img_view_transformer
    ref = dict(ref.to_dict() if hasattr(ref, 'to_dict') else ref)
    assert ref.pop('type') == 'MGHS_Stereo'
    norm = lambda v: {k: norm(x) for k, x in v.items()} if isinstance(v, dict) else \
        ([norm(x) for x in v] if isinstance(v, (list, tuple)) else v)
    assert norm(ref) == norm(synth.DHD_L_VIEW_TRANSFORMER)
    assert cfg.model.img_backbone.embed_dims == synth.DHD_L_STEREO_CHANNELS


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_mghs_stereo_matches_reference_construction():
    """MGHS_Stereo built from the DHD-L kwargs: parameter names / shapes of the stereo DepthNet (cost_volumn_net, the
    first block's downsample) and the 1/4-resolution frustum template equal the reference's, and that template is
    separable into the three axis vectors the cost-volume kernel reads."""
    from dhd_b200 import stereo as S
    from dhd_b200 import synth
    import projects.mmdet3d_plugin.models.necks.lss_heightmap as LH
    ns = ref_loader.load_reference()
    kw = dict(synth.DHD_L_VIEW_TRANSFORMER)
    ours, ref = LH.MGHS_Stereo(**kw), ns.MGHS_Stereo(**kw)
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert list(rs.keys()) == list(os_.keys())
    for k in rs:
        assert rs[k].shape == os_[k].shape, k
    assert any(k.startswith('depth_net.cost_volumn_net.') for k in os_) and 'depth_net.depth_conv.0.downsample.weight' in os_
    assert torch.equal(ours.cv_frustum, ref.cv_frustum) and ours.depth_net.bias == ref.depth_net.bias == 5.0
    fu, fv, fd = S.frustum_axes(ref.cv_frustum)
    D, H, W, _ = ref.cv_frustum.shape
    rebuilt = torch.stack((fu.view(1, 1, W).expand(D, H, W), fv.view(1, H, 1).expand(D, H, W),
                           fd.view(D, 1, 1).expand(D, H, W)), -1)
    assert torch.equal(rebuilt, ref.cv_frustum)


def test_stereo_host_side_refuses_cpu_tensors():
    """No CPU fallback on the stereo path either."""
    from dhd_b200 import stereo as S
    z = torch.zeros(1, 4, 4, 8)
    with pytest.raises(RuntimeError, match='CUDA'):
        S.cost_volume(z, z, 4, (16, 16), grid=torch.zeros(1, 16, 4, 2))
    with pytest.raises(RuntimeError, match='CUDA'):
        S.to_nhwc(torch.zeros(1, 8, 4, 4))
    with pytest.raises(RuntimeError, match='CUDA'):
        S.camera_table(torch.eye(4).view(1, 1, 4, 4), torch.eye(3).view(1, 1, 3, 3), torch.eye(3).view(1, 1, 3, 3),
                       torch.zeros(1, 1, 3))


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_plugin_depth_and_height_loss_equals_the_reference():
    """MGHS_Depth.get_depth_and_height_loss (LH:859-897) and the GT down-sampling it uses: plugin == unmodified reference
    on the same sparse maps (this is what the CUDA loss kernels are then compared with on the GPU)."""
    from dhd_b200 import synth
    import projects.mmdet3d_plugin.models.necks.lss_heightmap as LH
    ns = ref_loader.load_reference()
    kw = dict(synth.DHD_L_VIEW_TRANSFORMER, in_channels=64)
    kw['depthnet_cfg'] = dict(use_dcn=False, aspp_mid_channels=32, stereo=True, bias=5.)
    kw['heightnet_cfg'] = dict(use_dcn=False, aspp_mid_channels=32)
    ours, ref = LH.MGHS_Stereo(**kw), ns.MGHS_Stereo(**kw)
    g = torch.Generator().manual_seed(2)
    B, N, H, W = 1, 6, 512, 1408
    hit = torch.rand(B, N, H, W, generator=g) < 0.02
    gt_d = torch.where(hit, 0.2 + 60.0 * torch.rand(B, N, H, W, generator=g), torch.zeros(()))
    gt_h = torch.where(hit, -2.0 + 8.5 * torch.rand(B, N, H, W, generator=g), torch.zeros(()))
    assert torch.equal(ours.get_downsampled_gt_depth(gt_d), ref.get_downsampled_gt_depth(gt_d))
    assert torch.equal(ours.get_downsampled_gt_height(gt_h), ref.get_downsampled_gt_height(gt_h))
    d = torch.randn(B * N, ours.D, 32, 88, generator=g).softmax(1)
    h = torch.randn(B * N, ours.H, 32, 88, generator=g).softmax(1)
    for a, b in zip(ours.get_depth_and_height_loss(gt_d, gt_h, d, h), ref.get_depth_and_height_loss(gt_d, gt_h, d, h)):
        assert torch.allclose(a, b, rtol=1e-6)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
@pytest.mark.parametrize('name', ['DHD-M.py', 'DHD-L.py'])
def test_stereo_configs_build_the_hot_path_modules(name):
    """projects/configs/DHD/DHD-M.py / DHD-L.py, unchanged, through build_model: the DHD_stereo shell with the view
    transformer (MGHS_Stereo + stereo DepthNet), both pre-process nets, the BEV / voxel encoders, SFA and the head,
    under the reference's child names."""
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200 import compat as C
    cfg = C.Config.fromfile(os.path.join(ref_loader.load_reference_configs(), name))
    model = C.build_model(cfg.model)
    assert type(model).__name__ == 'DHD_stereo'
    vt = model.img_view_transformer
    assert type(vt).__name__ == 'MGHS_Stereo' and vt.depth_net.stereo and vt.D == 88 and vt.H == 65
    assert model.num_frame == 3 and model.temporal_frame == 2 and model.extra_ref_frames == 1
    names = [n for n, _ in model.named_children()]
    for want in ('img_view_transformer', 'img_bev_encoder_backbone', 'img_bev_encoder_neck', 'pre_process_net',
                 'pre_process_net_3d', 'img_voxel_encoder0', 'img_voxel_neck0', 'img_voxel_encoder1', 'img_voxel_encoder2',
                 'mix', 'occ_head'):
        assert want in names and getattr(model, want) is not None, want
    sd = model.state_dict()
    for key in ('pre_process_net.layers.0.0.conv1.weight', 'pre_process_net_3d.layers.0.0.downsample.weight',
                'img_view_transformer.depth_net.cost_volumn_net.0.weight', 'img_voxel_encoder2.inc.double_conv.0.weight',
                'mix.mysk_7.fc.0.weight', 'occ_head.final_conv.conv.weight'):
        assert key in sd, key
    assert sd['pre_process_net_3d.layers.0.0.conv1.weight'].shape[:2] == (1024, 1024)
    # the bev encoder sees both temporal frames on the channel axis (DHD_model.py:517)
    first = 'img_bev_encoder_backbone.' + ('layers.0.0.conv1.weight' if name == 'DHD-L.py' else 'inc.double_conv.0.weight')
    assert sd[first].shape[1] == 64 * 2            # CustomResNet in DHD-L, a UNet in DHD-M


def test_stereo_detector_frame_fusion_glue():
    """DHD_stereo.fuse_frames / pre-process restore (DHD_model.py:358-368, 517-541) with stand-in children: channel
    order of the collapsed z axis, the 4 / 4 / 8 plane split, and that chunk + stack undoes the collapse."""
    from projects.mmdet3d_plugin.models.detectors.DHD_model import DHD_stereo
    m = DHD_stereo.__new__(DHD_stereo)
    torch.nn.Module.__init__(m)
    seen = {}
    m._encode = lambda backbone, neck, x: x
    m.img_bev_encoder_backbone = m.img_bev_encoder_neck = None
    m.bev_encoder = lambda x: seen.setdefault('bev', x)
    m.voxel_encoder = lambda i, x: seen.setdefault('vox%d' % i, x)
    g = torch.Generator().manual_seed(0)
    B, C, Dy, Dx = 2, 3, 5, 4
    f2 = [torch.randn(B, C, 1, Dy, Dx, generator=g) for _ in range(2)]
    f3 = [torch.randn(B, C, 16, Dy, Dx, generator=g) for _ in range(2)]
    x_2d, x_3d = m.fuse_frames(f2, f3)
    cat3 = torch.cat(f3, dim=1)                                        # (B, 2C, 16, Dy, Dx)
    assert torch.equal(x_2d, torch.cat(f2, dim=1)[:, :, 0])
    want0 = torch.cat([cat3[:, :, z] for z in range(0, 4)], dim=1)
    want2 = torch.cat([cat3[:, :, z] for z in range(8, 16)], dim=1)
    assert torch.equal(seen['vox0'], want0) and torch.equal(seen['vox2'], want2)
    assert x_3d.shape == (B, 2 * C * 16, Dy, Dx) and torch.equal(x_3d[:, :2 * C * 4], want0)
    col = DHD_stereo._collapse_z(f3[0])                                # channel = z * C + c
    assert torch.equal(col[:, 5 * C + 1], f3[0][:, 1, 5])
    assert torch.equal(torch.stack(torch.chunk(col, 16, dim=1), dim=2), f3[0])


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_plugin_predictor_loss_equals_the_reference():
    """predictor.loss (occ_head.py:102-139) of the plugin -- CrossEntropyLoss from the LOSSES registry + the vectorised
    sem_scal / geo_scal terms -- against the UNMODIFIED reference predictor.loss driving the reference's own
    cross_entropy_loss.py and semkitti_loss.py: the three values and the gradient at the logits.  The one mmdet helper
    the reference file imports (mmdet 2.25.1 `weight_reduce_loss`, not in the reference tree) is restated here."""
    import importlib.util
    import sys
    import types
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
    ns = ref_loader.load_reference()

    def weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):      # mmdet/models/losses/utils.py
        if weight is not None:
            loss = loss * weight
        if avg_factor is None:
            return loss.mean() if reduction == 'mean' else (loss.sum() if reduction == 'sum' else loss)
        if reduction == 'mean':
            return loss.sum() / (avg_factor + torch.finfo(torch.float32).eps)
        if reduction != 'none':
            raise ValueError('avg_factor can not be used with reduction="sum"')
        return loss

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls
    saved = {k: sys.modules.get(k) for k in ('mmdet.models.builder', 'mmdet.models.losses', 'mmdet.models.losses.utils')}
    sys.modules['mmdet.models.builder'] = types.SimpleNamespace(LOSSES=_Reg())
    sys.modules['mmdet.models.losses'] = types.ModuleType('mmdet.models.losses')
    sys.modules['mmdet.models.losses.utils'] = types.SimpleNamespace(weight_reduce_loss=weight_reduce_loss)
    try:
        path = os.path.join(ref_loader.REF_ROOT, 'projects', 'mmdet3d_plugin', 'models', 'losses', 'cross_entropy_loss.py')
        spec = importlib.util.spec_from_file_location('ref_cross_entropy_loss', path)
        ce = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ce)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    kw = dict(in_dim=32, out_dim=32, Dz=16, use_mask=True, num_classes=18, use_predicter=True, class_balance=True,
              weight_ce=10.0, weight_geo=0.2, weight_sem=0.2)
    loss_cfg = dict(type='CrossEntropyLoss', use_sigmoid=False, ignore_index=255, loss_weight=1.0)
    ours = predictor(loss_occ=loss_cfg, **kw)
    ref = ns.predictor(loss_occ=loss_cfg, **kw)
    ref.loss_occ = ce.CrossEntropyLoss(use_sigmoid=False, ignore_index=255, loss_weight=1.0, class_weight=ref.cls_weights)
    assert type(ours.loss_occ).__name__ == 'CrossEntropyLoss' and ours.loss_occ.ignore_index == 255
    g = torch.Generator().manual_seed(6)
    B, Dx, Dy, Dz = 1, 20, 20, 16
    labels = torch.randint(0, 18, (B, Dx, Dy, Dz), generator=g)
    labels[labels == 5] = 6                              # a class absent from the target
    labels[0, 0, 0, :4] = 255                            # ignored voxels
    mask = torch.rand(B, Dx, Dy, Dz, generator=g) < 0.6
    for scale in (1.0, 5.0):
        logits = torch.randn(B, Dx, Dy, Dz, 18, generator=g) * scale
        a, b = logits.clone().requires_grad_(), logits.clone().requires_grad_()
        la, lb = ref.loss(a, labels, mask), ours.loss(b, labels, mask)
        assert set(la) == set(lb) == {'loss_occ', 'loss_voxel_sem_scal', 'loss_voxel_geo_scal'}
        for k in la:
            assert torch.allclose(la[k].double(), lb[k].double(), rtol=1e-5, atol=1e-6), (k, float(la[k]), float(lb[k]))
        sum(la.values()).backward()
        sum(lb.values()).backward()
        assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-8)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_plugin_mghs_exposes_every_reference_method():
    """Every public method / attribute of the reference's MGHS family exists on the plugin's (SURVEY 8(b) module API),
    and the two small helpers that are plain torch agree with the reference: downsample_sparse_map (LH:566-594) and
    init_acceleration_v2 (LH:234-258; same interval table, ranks equal as per-interval sets -- argsort is unstable)."""
    from dhd_b200 import synth
    import projects.mmdet3d_plugin.models.necks.lss_heightmap as LH
    ns = ref_loader.load_reference()
    kw = dict(synth.DHD_L_VIEW_TRANSFORMER, in_channels=64)
    kw['depthnet_cfg'] = dict(use_dcn=False, aspp_mid_channels=32, stereo=True, bias=5.)
    kw['heightnet_cfg'] = dict(use_dcn=False, aspp_mid_channels=32)
    ours, ref = LH.MGHS_Stereo(**kw), ns.MGHS_Stereo(**kw)
    public = [n for n in dir(ref) if not n.startswith('_') and n not in dir(torch.nn.Module)]
    missing = [n for n in public if not hasattr(ours, n)]
    assert not missing, missing
    g = torch.Generator().manual_seed(1)
    m = torch.where(torch.rand(1, 2, 64, 96, generator=g) < 0.03, torch.rand(1, 2, 64, 96, generator=g) * 5 - 1, torch.zeros(()))
    assert torch.equal(ours.downsample_sparse_map(m), ref.downsample_sparse_map(m))
    coor = torch.rand(1, 2, 8, 4, 6, 3, generator=g) * torch.tensor([90.0, 90.0, 8.0]) - torch.tensor([45.0, 45.0, 2.0])
    ours.init_acceleration_v2(coor)
    ref.init_acceleration_v2(coor)
    assert torch.equal(ours.interval_starts, ref.interval_starts) and torch.equal(ours.interval_lengths, ref.interval_lengths)
    assert torch.equal(ours.ranks_bev, ref.ranks_bev)
    for s, l in zip(ref.interval_starts.tolist(), ref.interval_lengths.tolist()):
        assert sorted(ours.ranks_depth[s:s + l].tolist()) == sorted(ref.ranks_depth[s:s + l].tolist())


def test_plugin_gen_grid_reproduces_the_reference_fixture():
    """DepthNet.gen_grid of the plugin (torch form of the reference API, depthnet.py:245-308) is bit-equal to the grid
    the unmodified reference produced for the fixture rig (tests/golden/depthnet.npz)."""
    from oracle import make_golden_depthnet as MGD
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.model_utils.depthnet import DepthNet, HeightNet
    gold = np.load(os.path.join(os.path.dirname(GOLD), 'depthnet.npz'))
    x, mlp, prev, curr = MGD.inputs()
    m = MGD.stereo_metas(prev, curr)
    H, W = MGD.INPUT
    for net in (DepthNet(64, 64, 32, MGD.n_depth(), stereo=True, use_dcn=False, bias=MGD.BIAS), HeightNet(64, 64, 65)):
        grid = net.gen_grid(m, MGD.B, MGD.NCAM, MGD.n_depth(), H // 4, W // 4, H, W)
        assert np.array_equal(grid.numpy(), gold['grid'])


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_plugin_classes_expose_the_reference_methods():
    """Public methods / attributes of every hot-path class of the reference exist on the plugin's class of that name."""
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.backbones.resnet import CustomResNet
    from projects.mmdet3d_plugin.models.backbones.unet import UNet
    from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
    from projects.mmdet3d_plugin.models.model_utils.depthnet import DepthNet, HeightNet
    from projects.mmdet3d_plugin.models.necks.lss_fpn import FPN_LSS
    from projects.mmdet3d_plugin.models.necks.mix import SFA
    ns = ref_loader.load_reference()
    pairs = [(SFA(512, 256), ns.SFA(512, 256)), (predictor(loss_occ=None), ns.predictor(loss_occ=None)),
             (HeightNet(64, 64, 65), ns.HeightNet(64, 64, 65)),
             (DepthNet(64, 64, 32, 88, stereo=True, use_dcn=False), ns.DepthNet(64, 64, 32, 88, stereo=True, use_dcn=False)),
             (UNet(64, 64), ns.UNet(64, 64)), (CustomResNet(64), ns.CustomResNet(64)), (FPN_LSS(640, 256), ns.FPN_LSS(640, 256))]
    base = set(dir(torch.nn.Module))
    for ours, ref in pairs:
        missing = [n for n in dir(ref) if not n.startswith('_') and n not in base and not hasattr(ours, n)]
        assert not missing, (type(ref).__name__, missing)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
@pytest.mark.parametrize('seed,bias,C,D', [(0, 0.0, 8, 5), (1, 5.0, 12, 9), (2, 2.5, 4, 33)])
def test_stereo_oracle_equals_live_reference_on_random_rigs(seed, bias, C, D):
    """oracle.stereo_sampling_grid / stereo_cost_volume (the checker of the CUDA kernel at full size) against the
    unmodified reference DepthNet.gen_grid / calculate_cost_volumn on random camera rigs, channel counts, depth counts
    and bias values -- including bias = 0 (the `== 0` test disabled) and points behind the previous camera."""
    ns = ref_loader.load_reference()
    g = torch.Generator().manual_seed(seed)
    B, N, H, W = 1, 3, 6, 10
    net = ns.DepthNet(16, 16, 8, D, use_dcn=False, aspp_mid_channels=8, stereo=True, bias=bias).eval()
    d = torch.linspace(0.5, 30.0, D).view(-1, 1, 1).expand(-1, H, W)
    u = torch.linspace(0, 4 * W - 1, W).view(1, 1, W).expand(D, H, W)
    v = torch.linspace(0, 4 * H - 1, H).view(1, H, 1).expand(D, H, W)
    frustum = torch.stack((u, v, d), -1)
    K = torch.tensor([[30.0, 0.0, 20.0], [0.0, 30.0, 12.0], [0.0, 0.0, 1.0]]).expand(B, N, 3, 3).contiguous()
    post_rots = torch.eye(3).expand(B, N, 3, 3).clone()
    post_rots[..., 0, 0] = post_rots[..., 1, 1] = 0.9 + 0.2 * torch.rand(B, N, generator=g)
    post_trans = torch.cat([4 * torch.rand(B, N, 2, generator=g) - 2, torch.zeros(B, N, 1)], -1)
    k2s = torch.eye(4).expand(B, N, 4, 4).clone()
    ang = 0.3 * (torch.rand(B, N, generator=g) - 0.5)
    k2s[..., 0, 0], k2s[..., 0, 2], k2s[..., 2, 0], k2s[..., 2, 2] = ang.cos(), ang.sin(), -ang.sin(), ang.cos()
    k2s[..., :3, 3] = torch.tensor([0.3, 0.0, 2.0]) * (torch.rand(B, N, 3, generator=g) - 0.2)   # some points end up behind
    prev, curr = torch.randn(B * N, C, H, W, generator=g), torch.randn(B * N, C, H, W, generator=g)
    prev[:, :, :2, :3] = 0.0
    metas = dict(k2s_sensor=k2s, intrins=K, post_rots=post_rots, post_trans=post_trans, frustum=frustum,
                 cv_downsample=4, downsample=16, grid_config=None, cv_feat_list=[prev, curr])
    with torch.no_grad():
        ref_grid = net.gen_grid(metas, B, N, D, H, W, 4 * H, 4 * W)
        ref_cv = net.calculate_cost_volumn(metas)
        grid = DO.stereo_sampling_grid(frustum, k2s, K, post_rots, post_trans, 4 * H, 4 * W)
        cv = DO.stereo_cost_volume(prev, curr, grid, D, bias)
    assert torch.equal(grid, ref_grid)
    assert torch.allclose(cv, ref_cv, rtol=1e-5, atol=1e-7)
    assert torch.allclose(cv.sum(1), torch.ones_like(cv[:, 0]), atol=1e-5)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_synth_dhd_s_model_cfg_is_the_reference_config():
    """dhd_b200.synth.dhd_s_model_cfg(images=True) -- what bench.py / scripts/bench_configs.py build on the GPU box, where
    the reference tree is absent -- is `model` of the unchanged projects/configs/DHD/DHD-S.py: every entry of ours equals
    the reference's (ours adds only the `precision` switch; entries ours leaves out are defaults / train_cfg)."""
    from dhd_b200 import compat as C
    from dhd_b200 import synth
    cfg = C.Config.fromfile(os.path.join(ref_loader.load_reference_configs(), 'DHD-S.py'))
    ref = cfg.model
    ref = ref.to_dict() if hasattr(ref, 'to_dict') else dict(ref)
    ours = synth.dhd_s_model_cfg('bf16', images=True)
    norm = lambda v: {k: norm(x) for k, x in v.items()} if isinstance(v, dict) else \
        ([norm(x) for x in v] if isinstance(v, (list, tuple)) else v)

    def check(a, b, path):
        """every key of ours (a) except `precision` exists in the reference (b) with the same value"""
        for k, v in a.items():
            if k == 'precision':
                continue
            assert k in b, path + k
            if isinstance(v, dict):
                check(v, norm(b[k]), path + k + '.')
            else:
                assert norm(v) == norm(b[k]), (path + k, v, b[k])
    check(norm(ours), norm(ref), 'model.')
    assert set(k for k in ref if k not in ours) <= {'train_cfg', 'test_cfg', 'pretrained', 'upsample'}


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_synth_dhd_l_model_cfg_is_the_reference_config():
    """dhd_b200.synth.dhd_l_model_cfg() (scripts/bench_configs.py dhd_l, BASELINE configs[4]) against `model` of the
    unchanged projects/configs/DHD/DHD-L.py: every entry of ours equals the reference's; ours leaves out the image
    backbone / neck (Swin-B + FPN_LSS) and adds the `precision` switch."""
    from dhd_b200 import compat as C
    from dhd_b200 import synth
    cfg = C.Config.fromfile(os.path.join(ref_loader.load_reference_configs(), 'DHD-L.py'))
    ref = cfg.model
    ref = ref.to_dict() if hasattr(ref, 'to_dict') else dict(ref)
    ours = synth.dhd_l_model_cfg('bf16')
    norm = lambda v: {k: norm(x) for k, x in v.items()} if isinstance(v, dict) else \
        ([norm(x) for x in v] if isinstance(v, (list, tuple)) else v)

    def check(a, b, path):
        for k, v in a.items():
            if k == 'precision':
                continue
            assert k in b, path + k
            if isinstance(v, dict):
                check(v, norm(b[k]), path + k + '.')
            else:
                assert norm(v) == norm(b[k]), (path + k, v, b[k])
    check(norm(ours), norm(ref), 'model.')
    assert set(k for k in ref if k not in ours) <= {'img_backbone', 'img_neck', 'train_cfg', 'test_cfg', 'pretrained', 'upsample',
                                                    'with_prev'}
