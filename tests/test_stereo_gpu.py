"""GPU parity tests of the camera-aware DepthNet (DHD-M / DHD-L) and its plane-sweep stereo cost volume
(dhd_b200/csrc/stereo.cu) against (a) tests/golden/depthnet.npz, made by running the UNMODIFIED reference DepthNet
(oracle/make_golden_depthnet.py), and (b) the torch restatement in oracle/dense_oracle.py run on the GPU at DHD-L size.

Tolerances.  Given the sampling grid, the cost volume is fp32 arithmetic in a different summation order: probabilities
within rtol 1e-4 + atol 1e-6.  With the geometry fused into the kernel, the sampling coordinate itself is only defined
to fp32 rounding of a 3-matrix chain (ulp(ix) = 3e-5 px at ix ~ 350); the coordinates must agree to 2e-5 of the
[-1, 1] range and the probabilities to rtol 2e-3.  DepthNet logits: atol 1e-4 + rtol 1e-4 (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import dense_oracle as DO
from oracle import make_golden_depthnet as MG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'depthnet.npz')


def close(got, ref, what, atol, rtol):
    got, ref = got.detach().float().cpu(), torch.as_tensor(ref).float().cpu()
    assert got.shape == ref.shape, '%s: shape %s vs %s' % (what, tuple(got.shape), tuple(ref.shape))
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    assert not bad.any(), '%s: %d / %d elements off, max abs err %.3g (ref scale %.3g)' % (
        what, int(bad.sum()), bad.numel(), err.max(), ref.abs().max())


def metas_cuda(prev, curr):
    m = MG.stereo_metas(prev, curr)
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in m.items()} | \
        {'cv_feat_list': [None if prev is None else prev.cuda(), curr.cuda()]}


def test_cost_volume_given_the_reference_grid(cuda_lib):
    from dhd_b200 import stereo as S
    gold = np.load(GOLD)
    _, _, prev, curr = MG.inputs()
    H, W = MG.INPUT
    grid = torch.from_numpy(gold['grid']).cuda()
    p, c = S.to_nhwc(prev.cuda()), S.to_nhwc(curr.cuda())
    assert torch.equal(p.cpu(), prev.permute(0, 2, 3, 1)) and torch.equal(c.cpu(), curr.permute(0, 2, 3, 1))
    cv, _ = S.cost_volume(p, c, MG.n_depth(), (H, W), bias=MG.BIAS, grid=grid)
    close(cv, gold['cost_volume'], 'cost volume (grid given)', 1e-6, 1e-4)
    assert torch.allclose(cv.sum(1), torch.ones_like(cv[:, 0]), atol=1e-5)
    # without the bias the volume differs where samples fall outside: the `== 0` test is live in this fixture
    cv0, _ = S.cost_volume(p, c, MG.n_depth(), (H, W), bias=0.0, grid=grid)
    want0 = DO.stereo_cost_volume(prev, curr, torch.from_numpy(gold['grid']), MG.n_depth(), 0.0)
    close(cv0, want0, 'cost volume (no bias)', 1e-6, 1e-4)
    assert (cv0 - cv).abs().max() > 1e-3


def test_cost_volume_fused_geometry(cuda_lib):
    from dhd_b200 import stereo as S
    gold = np.load(GOLD)
    _, _, prev, curr = MG.inputs()
    H, W = MG.INPUT
    m = metas_cuda(prev, curr)
    cam = S.camera_table(m['k2s_sensor'], m['intrins'], m['post_rots'], m['post_trans'])
    p, c = S.to_nhwc(prev.cuda()), S.to_nhwc(curr.cuda())
    cv, grid = S.cost_volume(p, c, MG.n_depth(), (H, W), bias=MG.BIAS, frustum=m['frustum'], cam=cam, want_grid=True)
    ref_grid = torch.from_numpy(gold['grid'])
    gerr = (grid.cpu() - ref_grid).abs().max().item()
    print('sampling grid: max |diff| vs the reference = %.3g (bit-equal: %s)' % (gerr, torch.equal(grid.cpu(), ref_grid)))
    assert gerr <= 2e-5
    assert torch.equal(grid.cpu() == -2, ref_grid == -2)          # the same points are behind the previous camera
    close(cv, gold['cost_volume'], 'cost volume (fused geometry)', 1e-6, 2e-3)


def test_cost_volume_bf16_features_and_activation_output(cuda_lib):
    from dhd_b200 import dense as D
    from dhd_b200 import stereo as S
    gold = np.load(GOLD)
    _, _, prev, curr = MG.inputs()
    H, W = MG.INPUT
    grid = torch.from_numpy(gold['grid']).cuda()
    pb, cb = S.to_nhwc(prev.cuda(), bf16=True), S.to_nhwc(curr.cuda(), bf16=True)
    # exact restatement on the rounded features: the kernel's bf16 mode only changes the operand type
    want = DO.stereo_cost_volume(prev.bfloat16().float(), curr.bfloat16().float(), torch.from_numpy(gold['grid']),
                                 MG.n_depth(), MG.BIAS)
    act = D.Act.empty(pb.shape[0], pb.shape[1], pb.shape[2], 64, 3, 'cuda')
    act.data.fill_(7.0)                                            # the kernel must overwrite the padding channels
    cv, _ = S.cost_volume(pb, cb, MG.n_depth(), (H, W), bias=MG.BIAS, grid=grid,
                          out=torch.empty(pb.shape[0], MG.n_depth(), pb.shape[1], pb.shape[2], device='cuda'), out_act=act)
    close(cv, want, 'cost volume (bf16 features)', 1e-6, 1e-4)
    full = act.float()                                             # (BN, 64, H, W) sum of the three bf16 parts
    close(full[:, :MG.n_depth()], cv, 'split-bf16 activation', 1e-7, 1e-6)
    assert float(full[:, MG.n_depth():].abs().max()) == 0.0


def build_nets(precision='fp32'):
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.model_utils.depthnet import DepthNet
    D = MG.n_depth()
    mono = DepthNet(MG.C_IN, MG.C_MID_MONO, MG.C_CTX, D, use_dcn=True, use_aspp=True, precision=precision).eval()
    st = DepthNet(MG.C_IN, MG.C_IN, MG.C_CTX, D, use_dcn=False, aspp_mid_channels=32, stereo=True, bias=MG.BIAS,
                  precision=precision).eval()
    sdm, sds = DO.seeded_state_dict(mono, 41), DO.seeded_state_dict(st, 42)
    gold = np.load(GOLD)
    assert MG.sha_sd(sdm) == str(gold['sha_mono']) and MG.sha_sd(sds) == str(gold['sha_stereo'])
    mono.load_state_dict(sdm)
    st.load_state_dict(sds)
    return mono.cuda(), st.cuda()


def test_depthnet_mono_matches_reference_fixture(cuda_lib):
    gold = np.load(GOLD)
    mono, st = build_nets()
    x, mlp, prev, curr = MG.inputs()
    y = mono(x.cuda(), mlp.cuda())
    close(y, gold['mono'], 'DepthNet (DHD-M form)', 1e-4, 1e-4)
    with pytest.raises(RuntimeError, match='stereo'):
        st(x.cuda(), mlp.cuda())                                   # stereo=True needs stereo_metas, as in the reference


def test_depthnet_stereo_matches_reference_fixture(cuda_lib):
    gold = np.load(GOLD)
    _, st = build_nets()
    x, mlp, prev, curr = MG.inputs()
    cv = st.calculate_cost_volumn(metas_cuda(prev, curr))
    close(cv, gold['cost_volume'], 'DepthNet.calculate_cost_volumn', 1e-6, 2e-3)
    y = st(x.cuda(), mlp.cuda(), metas_cuda(prev, curr))
    close(y, gold['stereo'], 'DepthNet (DHD-L form, stereo)', 1e-4, 1e-4)
    y0 = st(x.cuda(), mlp.cuda(), metas_cuda(None, curr))
    close(y0, gold['stereo_first'], 'DepthNet (stereo, no previous frame)', 1e-4, 1e-4)
    assert (y - y0).abs().max() > 1e-2                             # the cost volume does reach the logits


def test_depthnet_bf16_speed_mode_is_close(cuda_lib):
    gold = np.load(GOLD)
    _, st = build_nets('bf16')
    x, mlp, prev, curr = MG.inputs()
    y = st(x.cuda(), mlp.cuda(), metas_cuda(prev, curr)).cpu()
    ref = torch.from_numpy(gold['stereo'])
    assert (y - ref).abs().max() < 3e-2 * ref.abs().max()


def dhdl_case(BN, C, seed=3):
    """DHD-L stereo geometry (512x1408 input, 1/4 map 128x352, D = 88) for BN images of the synthetic rig."""
    import math
    H, W, D = 128, 352, 88
    g = torch.Generator().manual_seed(seed)
    k = torch.ones(1, 1, 5, 5) / 25.0
    def feat():
        x = torch.randn(BN, C, H, W, generator=g)
        return torch.nn.functional.conv2d(x.flatten(0, 1)[:, None], k, padding=2).view(BN, C, H, W).cuda()
    d = torch.arange(1.0, 45.0, 0.5).view(-1, 1, 1).expand(-1, H, W)
    u = torch.linspace(0, 4 * W - 1, W).view(1, 1, W).expand(D, H, W)
    v = torch.linspace(0, 4 * H - 1, H).view(1, H, 1).expand(D, H, W)
    frustum = torch.stack((u, v, d), -1).cuda()
    s = 4 * W / 1600.0
    intr = torch.tensor([[1266.0, 0.0, 816.0], [0.0, 1266.0, 491.0], [0.0, 0.0, 1.0]]).expand(1, BN, 3, 3).contiguous()
    post_rots = torch.diag(torch.tensor([s, s, 1.0])).expand(1, BN, 3, 3).contiguous()
    post_trans = torch.tensor([0.0, -140.0 * s, 0.0]).expand(1, BN, 3).contiguous()
    k2s = torch.eye(4).expand(1, BN, 4, 4).contiguous()
    return feat, frustum, dict(k2s_sensor=k2s.cuda(), intrins=intr.cuda(), post_rots=post_rots.cuda(),
                               post_trans=post_trans.cuda()), (H, W, D)


def test_full_size_properties_and_gpu_oracle(cuda_lib):
    """DHD-L size: (1) identical frames + identity ego motion => every hypothesis matches => uniform 1/D;
    (2) a moving rig against the torch restatement of the reference run on the GPU."""
    from dhd_b200 import stereo as S
    BN, C = 2, 128
    feat, frustum, cams, (H, W, D) = dhdl_case(BN, C)
    f = feat()
    cam = S.camera_table(**cams)
    fn = S.to_nhwc(f)
    cv, _ = S.cost_volume(fn, fn, D, (4 * H, 4 * W), bias=5.0, frustum=frustum, cam=cam)
    assert torch.allclose(cv.sum(1), torch.ones_like(cv[:, 0]), atol=1e-5)
    assert (cv - 1.0 / D).abs().max() < 0.05 / D
    # moving rig: 0.8 m forward, small yaw
    import math
    a = math.radians(1.5)
    k2s = torch.eye(4)
    k2s[:3, :3] = torch.tensor([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]])
    k2s[:3, 3] = torch.tensor([0.05, 0.0, 0.8])
    cams['k2s_sensor'] = k2s.expand(1, BN, 4, 4).contiguous().cuda()
    prev = feat()
    cam = S.camera_table(**cams)
    cv, grid = S.cost_volume(S.to_nhwc(prev), fn, D, (4 * H, 4 * W), bias=5.0, frustum=frustum, cam=cam, want_grid=True)
    ref_grid = DO.stereo_sampling_grid(frustum, cams['k2s_sensor'], cams['intrins'], cams['post_rots'],
                                       cams['post_trans'], 4 * H, 4 * W)
    assert (grid - ref_grid).abs().max() <= 2e-5
    want = DO.stereo_cost_volume(prev, f, grid, D, 5.0)          # sampler + cost + softmax on the kernel's own grid
    close(cv, want, 'cost volume at DHD-L size vs torch on the GPU', 1e-6, 2e-4)
    frac_inside = float(((grid.abs() <= 1).all(-1)).float().mean())
    assert 0.5 < frac_inside < 1.0


def test_mghs_stereo_dhdl_full_size(cuda_lib):
    """BASELINE configs[4]: the plugin's MGHS_Stereo with the DHD-L.py kwargs at full size (B=1: 6 cameras, 512x1408,
    C_in=512, D=88, stereo features 128 ch @128x352).  Depth / context of the stereo DepthNet against the torch
    restatement of the reference run on the GPU in fp32 (TF32 off) with the same seeded weights; output shapes of the
    collapse_z=False pool (its values are pinned at this size by tests/test_pool_gpu.py)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'scripts'))
    import bench_dhdl
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    vt, args, metas = bench_dhdl.build('fp32', 1)
    sd = {k: v.cuda() for k, v in DO.seeded_state_dict(vt, 77).items()}
    vt.load_state_dict(sd)
    with torch.no_grad():
        bev, bev_z, depth, height = vt(args, metas)
        x = args[0].flatten(0, 1)
        prev, curr = metas['cv_feat_list']
        grid = DO.stereo_sampling_grid(metas['frustum'], metas['k2s_sensor'], metas['intrins'], metas['post_rots'],
                                       metas['post_trans'], 512, 1408)
        cv = DO.stereo_cost_volume(prev, curr, grid, vt.D, 5.0)
        dsd = {k[len('depth_net.'):]: v for k, v in sd.items() if k.startswith('depth_net.')}
        y = DO.depthnet_forward(dsd, x, args[7], cv)
        want_depth = y[:, :vt.D].softmax(1)
        h = DO.heightnet_forward(sd, x, args[7], prefix='height_net.').softmax(1)
    assert bev.shape == (1, 64, 1, 200, 200) and bev_z.shape == (1, 64, 16, 200, 200)
    assert torch.isfinite(bev).all() and torch.isfinite(bev_z).all() and float(bev.abs().sum()) > 0
    close(depth, want_depth, 'DHD-L stereo depth distribution', 1e-4, 1e-3)
    close(height, h, 'DHD-L height distribution', 1e-4, 1e-3)


def test_dhd_stereo_detector_end_to_end(cuda_lib):
    """The DHD_stereo shell with the DHD-L wiring (key + previous frame + stereo reference frame, both pre-process
    nets, CustomResNet + FPN_LSS and three UNets at the DHD-L widths, SFA, predictor) from per-frame image features
    to occupancy classes, at a reduced image size: shapes, finiteness, a non-degenerate class map."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'scripts'))
    import run_dhd_stereo
    res = run_dhd_stereo.run()
    assert res['occ'] == [1, 200, 200, 16, 18] and res['depth'] == [6, 88, 4, 11] and res['height'] == [6, 65, 4, 11]
    assert res['finite'] and res['occ_abs_mean'] > 0 and res['classes_found'] > 1
