"""GPU parity tests of the dense hot-path modules (plugin API -> engines -> C-ABI -> tcgen05)
against (a) the fixture generated from the REAL reference modules and (b) the CPU oracle, on the
same seeded weights and inputs.

Tolerance: north_star asks occupancy logits within 1e-4 in fp32.  In the 'fp32' precision mode
(6-term split-bf16 MMAs, fp32 accumulation) every module output must be within
atol 1e-4 + rtol 1e-4 of the reference; the 'bf16' speed mode is checked against its own, looser
bound (2e-2 of the output scale) and reported separately."""
import os

import numpy as np
import pytest
import torch

from oracle import dense_oracle as DO
from oracle import make_golden_dense as MG
from oracle import mghs_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'dense_modules.npz')
ATOL, RTOL = 1e-4, 1e-4


def build(precision):
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
    from projects.mmdet3d_plugin.models.model_utils.depthnet import HeightNet
    from projects.mmdet3d_plugin.models.necks.mix import SFA
    hn = HeightNet(256, 256, 65, precision=precision).eval()
    sfa = SFA(512, 256, precision=precision).eval()
    head = predictor(in_dim=256, out_dim=256, Dz=16, num_classes=18, use_predicter=True,
                     class_balance=False, loss_occ=None, precision=precision).eval()
    dn = torch.nn.Conv2d(256, 108, 1)
    for m, s in zip((hn, sfa, head, dn), (21, 22, 23, 24)):
        m.load_state_dict(DO.seeded_state_dict(m, s))
    return hn.cuda(), sfa.cuda(), head.cuda(), dn.cuda()


def close(got, ref, what, atol=ATOL, rtol=RTOL):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, '%s: shape %s vs %s' % (what, tuple(got.shape), tuple(ref.shape))
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    assert not bad.any(), '%s: %d / %d elements off, max abs err %.3g (ref scale %.3g)' % (
        what, int(bad.sum()), bad.numel(), err.max(), ref.abs().max())


def test_heightnet_matches_reference_fixture(cuda_lib):
    gold = np.load(GOLD)
    hn, _, _, _ = build('fp32')
    x, mlp, _ = MG.inputs()
    height = hn(x.cuda(), mlp.cuda())
    close(height, torch.from_numpy(gold['height']), 'HeightNet logits')
    sm = hn(x.cuda(), mlp.cuda(), softmax=True)
    close(sm, torch.from_numpy(gold['height']).softmax(1), 'HeightNet softmax', atol=1e-5)


def test_heightnet_without_dcn_matches_oracle(cuda_lib):
    """DHD-L switches the deformable conv off (use_dcn=False, DHD-L.py:118-119)."""
    from projects.mmdet3d_plugin.models.model_utils.depthnet import HeightNet
    hn = HeightNet(256, 256, 65, use_dcn=False).eval()
    hn.load_state_dict(DO.seeded_state_dict(hn, 31))
    x, mlp, _ = MG.inputs()
    with torch.no_grad():
        ref = DO.heightnet_forward(hn.state_dict(), x, mlp)
    close(hn.cuda()(x.cuda(), mlp.cuda()), ref, 'HeightNet(use_dcn=False)')


def test_depth_head_matches_reference_fixture(cuda_lib):
    from dhd_b200 import dense as D
    from dhd_b200.modules import DepthHeadEngine
    gold = np.load(GOLD)
    _, _, _, dn = build('fp32')
    x, _, _ = MG.inputs()
    depth, feat = DepthHeadEngine(dn, 44, 'fp32', 'cuda')(D.pack_input(x.cuda(), 3))
    y = torch.from_numpy(gold['depth_net'])
    close(depth, y[:, :44].softmax(1), 'depth softmax', atol=1e-5)
    close(feat.permute(0, 3, 1, 2), y[:, 44:], 'context feature')


def test_sfa_and_predictor_match_reference_fixture(cuda_lib):
    gold = np.load(GOLD)
    _, sfa, head, _ = build('fp32')
    _, _, bev = MG.inputs()
    fused = sfa(bev.cuda())
    close(fused, torch.from_numpy(gold['sfa']), 'SFA output')
    occ = head(sfa(bev.cuda(), return_act=True))
    close(occ, torch.from_numpy(gold['occ']), 'occupancy logits')
    # channels_last input and the tensor (non-Act) path of the head
    fused_cl = sfa(bev.cuda().contiguous(memory_format=torch.channels_last))
    close(fused_cl, fused, "SFA channels_last vs NCHW input", atol=1e-5, rtol=1e-5)   # squeeze mean uses fp32 atomics
    close(head(fused), torch.from_numpy(gold['occ']), 'occupancy logits (tensor path)')


def test_bf16_speed_mode_is_close(cuda_lib):
    gold = np.load(GOLD)
    hn, sfa, head, _ = build('bf16')
    x, mlp, bev = MG.inputs()
    occ = head(sfa(bev.cuda(), return_act=True)).cpu()
    ref = torch.from_numpy(gold['occ'])
    assert (occ - ref).abs().max() <= 2e-2 * ref.abs().max()
    h = hn(x.cuda(), mlp.cuda()).cpu()
    ref = torch.from_numpy(gold['height'])
    assert (h - ref).abs().max() <= 3e-2 * ref.abs().max()


def test_conv_per_image_weights_and_bf16_residual(cuda_lib):
    """dhd_conv_desc.w_image_rows (image n convolves with its own weight rows) and res_b16 (bf16 identity path added
    before the activation) against torch on the same bf16-rounded operands."""
    from dhd_b200 import dense as D
    g = torch.Generator().manual_seed(5)
    N, H, W, Cin, Cout = 3, 20, 24, 128, 192
    x = torch.randn(N, Cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(N, Cout, Cin, generator=g) / Cin ** 0.5).bfloat16().float()
    r = torch.randn(N, Cout, H, W, generator=g).bfloat16().float()
    bias = torch.randn(Cout, generator=g)
    ref = torch.relu(torch.einsum('nchw,noc->nohw', x, w) + bias[None, :, None, None] + r)
    xa = D.pack_input(x.cuda(), 1)
    ra = D.pack_input(r.cuda(), 1)
    out = D.Act.empty(N, H, W, Cout, 1, 'cuda')
    wq = w.reshape(N * Cout, 1, 1, Cin).bfloat16().cuda().contiguous()
    D.conv2d(xa, wq, Cout, precision='bf16', bias=bias.cuda(), image_weights=True, residual_act=ra,
             segs=[dict(act='relu', out_act=out)])
    got = out.float().cpu()
    assert (got - ref).abs().max() <= 1e-2 * ref.abs().max() + 1e-3        # bf16 output rounding
    # 3x3 with a bf16 residual, shared weights
    w3 = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).bfloat16().float()
    ref3 = torch.nn.functional.conv2d(x, w3, padding=1) + r
    out3 = torch.empty(N, H, W, Cout, device='cuda')
    D.conv2d(xa, D.pack_weight(w3.cuda(), 1), Cout, ksize=3, precision='bf16', residual_act=ra,
             segs=[dict(out_f32=(out3, D.nhwc_strides(Cout, H, W)))])
    assert (out3.permute(0, 3, 1, 2).cpu() - ref3).abs().max() <= 2e-5 * ref3.abs().max() + 2e-5


@pytest.mark.parametrize('shape', [(2, 200, 200, 256, 256, 3, 1), (6, 16, 44, 256, 256, 3, 1), (1, 24, 40, 512, 320, 1, 1),
                                   (3, 50, 50, 128, 512, 3, 6), (1, 33, 21, 64, 160, 3, 1), (1, 8, 16, 64, 256, 1, 1)])
def test_conv_cta_pair_kernel_is_bit_equal_to_single_cta(cuda_lib, shape):
    """tcgen05.mma.cta_group::2 path (clusters of two CTAs, M = 256, half of the weight tile per CTA) == the one-CTA
    kernel, bit for bit: odd tile counts (a pair with a missing half), several N tiles, dilated taps that are dead for
    one tile of a pair only, a single tile, bias / ReLU / bf16 residual epilogues, and against torch."""
    from dhd_b200 import _lib
    from dhd_b200 import dense as D
    N, H, W, Cin, Cout, k, dil = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, Cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).bfloat16().float()
    r = torch.randn(N, Cout, H, W, generator=g).bfloat16().float()
    bias = torch.randn(Cout, generator=g)
    xa, ra, wq = D.pack_input(x.cuda(), 1), D.pack_input(r.cuda(), 1), D.pack_weight(w.cuda(), 1)
    lib = _lib.load()
    outs = []
    prev = lib.dhd_conv_pair_mode(-1)
    try:
        for mode in (2, 0):
            lib.dhd_conv_pair_mode(mode)
            o16 = D.Act.empty(N, H, W, (Cout + 63) // 64 * 64, 1, 'cuda')
            o16.data.zero_()
            o32 = torch.empty(N, H, W, Cout, device='cuda')
            D.conv2d(xa, wq, Cout, ksize=k, dilation=dil, precision='bf16', bias=bias.cuda(), residual_act=ra if Cout % 64 == 0 else None,
                     segs=[dict(act='relu', out_act=o16)])
            D.conv2d(xa, wq, Cout, ksize=k, dilation=dil, precision='bf16',
                     segs=[dict(out_f32=(o32, D.nhwc_strides(Cout, H, W)))])
            torch.cuda.synchronize()
            outs.append((o16.data.clone(), o32))
    finally:
        lib.dhd_conv_pair_mode(prev)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref = torch.nn.functional.conv2d(x.cuda(), w.cuda(), padding=dil * (k // 2), dilation=dil)
    err = (outs[0][1].permute(0, 3, 1, 2) - ref).abs().max()
    assert float(err) <= 1e-4 * float(ref.abs().max()) + 1e-5, float(err)


def test_conv_batch_launch_is_bit_equal_to_single_launches(cuda_lib):
    """dhd_conv2d_fwd_batch: four heterogeneous convolutions (1x1, dilated 3x3 with dead taps, a narrow softmax head
    with two segments, a different input tensor) in one persistent launch == the same layers launched one by one."""
    from dhd_b200 import dense as D
    g = torch.Generator().manual_seed(7)
    N, H, W = 3, 16, 44
    xa = D.pack_input(torch.randn(N, 256, H, W, generator=g).cuda(), 1)
    xb = D.pack_input(torch.randn(N, 128, H, W, generator=g).cuda(), 1)
    mk = lambda co, ci, k: D.pack_weight((torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5).cuda(), 1)
    w1, w2, w3, w4 = mk(256, 256, 1), mk(256, 256, 3), mk(108, 256, 1), mk(192, 128, 3)
    bias = torch.randn(256, generator=g).cuda()

    def layers(defer):
        o1 = D.Act.empty(N, H, W, 256, 1, 'cuda')
        o2 = D.Act.empty(N, H, W, 256, 1, 'cuda')
        dep, ctx = torch.empty(N, 44, H, W, device='cuda'), torch.empty(N, H, W, 64, device='cuda')
        o4 = torch.empty(N, H, W, 192, device='cuda')
        r = [D.conv2d(xa, w1, 256, precision='bf16', bias=bias, segs=[dict(act='relu', out_act=o1)], defer=defer),
             D.conv2d(xa, w2, 256, ksize=3, dilation=18, precision='bf16', segs=[dict(act='relu', out_act=o2)], defer=defer),
             D.conv2d(xa, w3, 108, precision='bf16', defer=defer,
                      segs=[dict(c_lo=0, c_hi=44, act='softmax', out_f32=(dep, D.nchw_strides(44, H, W))),
                            dict(c_lo=44, c_hi=108, out_f32=(ctx, D.nhwc_strides(64, H, W)))]),
             D.conv2d(xb, w4, 192, ksize=3, precision='bf16', defer=defer,
                      segs=[dict(out_f32=(o4, D.nhwc_strides(192, H, W)))])]
        return r, (o1.data, o2.data, dep, ctx, o4)
    _, single = layers(False)
    deferred, batched = layers(True)
    D.conv2d_batch(deferred)
    torch.cuda.synchronize()
    for a, b, what in zip(single, batched, ('1x1', 'dilated 3x3', 'depth softmax', 'context', '3x3 on another input')):
        assert torch.equal(a, b), what
    assert float(batched[2].sum(1).sub(1).abs().max()) < 1e-5


def test_sfa_fold_gate_and_bf16_blend_kernels(cuda_lib):
    """dhd_sfa_fold_gate / dhd_sfa_blend_b16 against their definitions (mix.py:41-57)."""
    import ctypes
    from dhd_b200 import _lib
    from dhd_b200 import dense as D
    g = torch.Generator().manual_seed(6)
    N, C, Co, HW = 2, 64, 48, 37 * 5
    w, a1 = torch.randn(Co, C, generator=g).cuda(), torch.rand(N, C, generator=g).cuda()
    out = torch.empty(N, Co, 2 * C, dtype=torch.bfloat16, device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.load().dhd_sfa_fold_gate(P(w), P(a1), N, Co, C, P(out), st), 'fold')
    want = torch.cat([w[None] * a1[:, None], w[None] * (1 - a1[:, None])], 2).bfloat16()
    assert torch.equal(out, want)
    x = torch.randn(N, 37, 5, 2 * C, generator=g).bfloat16().cuda()
    a2 = torch.rand(N, 37, 5, C, generator=g).bfloat16().cuda()
    res = torch.empty(N, 37, 5, C, dtype=torch.bfloat16, device='cuda')
    _lib.check(_lib.load().dhd_sfa_blend_b16(P(x), 2 * C, 0, C, N, HW, P(a1), P(a2), C, 0, P(res), C, 0, st), 'blend')
    bev, vox, g1, g2 = x[..., :C].float(), x[..., C:].float(), a1[:, None, None, :], a2.float()
    want = (g2 * (g1 * bev) + (1 - g2) * ((1 - g1) * vox)).bfloat16()
    assert torch.equal(res, want)


def test_bf16_lean_paths_match_layer_by_layer(cuda_lib, monkeypatch):
    """bf16 speed mode: SFA with the channel gate folded into per-image weights and bf16 gate / shortcut tensors, and
    HeightNet with bf16 identity paths, against the layer-by-layer form of the same mode (fp32 side tensors)."""
    hn, sfa, head, _ = build('bf16')
    x, mlp, bev = MG.inputs()
    outs = {}
    for lean in ('1', '0'):
        monkeypatch.setenv('DHD_SFA_LEAN', lean)
        monkeypatch.setenv('DHD_BF16_RESIDUAL', lean)
        outs[lean] = (sfa(bev.cuda()).cpu(), hn(x.cuda(), mlp.cuda()).cpu())
    for a, b, what in zip(outs['1'], outs['0'], ('SFA', 'HeightNet')):
        assert (a - b).abs().max() <= 1.5e-2 * b.abs().max(), (what, float((a - b).abs().max()), float(b.abs().max()))
    gold = np.load(GOLD)
    ref = torch.from_numpy(gold['sfa'])
    assert (outs['1'][0] - ref).abs().max() <= 2e-2 * ref.abs().max()


def test_mghs_forward_end_to_end(cuda_lib):
    """Plugin MGHS.forward (dense front + fused pool) on the MINI rig: dense outputs against the
    oracle, pooled BEV tensors against the oracle's view_transform fed with the SAME depth /
    context / height (so an argmax tie in the height head cannot flip a mask between the two)."""
    from projects.mmdet3d_plugin.models.necks.lss_heightmap import MGHS
    cfg, B = O.MINI, 2
    grids = cfg['mask_grids']
    vt = MGHS(grid_config=dict(cfg['bev_grid'], depth=cfg['depth']), input_size=cfg['input_size'],
              in_channels=256, out_channels=64, height_range=cfg['height_range'], height_interval=0.1,
              mask_range=cfg['mask_range'], mask_1_grid=dict(grids[0], depth=cfg['depth']),
              mask_2_grid=dict(grids[1], depth=cfg['depth']), mask_3_grid=dict(grids[2], depth=cfg['depth']),
              downsample=16).eval()
    # MINI uses coarse slabs on the x/y grid of the BEV pass for the fused path
    for g in (vt.mask_1_grid, vt.mask_2_grid, vt.mask_3_grid):
        g['x'], g['y'] = [-40, 40, 0.4], [-40, 40, 0.4]
    vt.load_state_dict(DO.seeded_state_dict(vt, 41))
    rig = O.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=3)
    s2e, e2g, K, pr, pt, bda = rig
    N, fH, fW = cfg['ncams'], 4, 11
    x = DO.seeded_tensor((B, N, 256, fH, fW), 42)
    mlp = vt.get_mlp_input(s2e, e2g, K, pr, pt, bda)
    sd = vt.state_dict()
    with torch.no_grad():
        d_ref, f_ref = DO.depth_head_forward(sd, x.view(B * N, 256, fH, fW), vt.D)
        h_ref = DO.heightnet_forward(sd, x.view(B * N, 256, fH, fW), mlp, prefix='height_net.').softmax(1)
    vt = vt.cuda()
    args = [t.cuda() for t in (x, s2e, e2g, K, pr, pt, bda, mlp)]
    bev, depth, height, lo, mid, hi = vt(args)
    close(depth, d_ref, 'depth', atol=1e-5)
    close(height, h_ref, 'height', atol=1e-5)
    assert bev.shape == (B, 64, 200, 200) and lo.shape == (B, 256, 200, 200) and hi.shape == (B, 512, 200, 200)
    # pool against the oracle on the module's own dense outputs
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample'])
    feat_nchw = vt._last_feat.permute(0, 3, 1, 2).contiguous().cpu() if hasattr(vt, '_last_feat') else None
    if feat_nchw is None:
        from dhd_b200 import dense as D
        from dhd_b200.modules import DepthHeadEngine
        _, f = DepthHeadEngine(vt.depth_net, vt.D, 'fp32', 'cuda')(D.pack_input(args[0].view(B * N, 256, fH, fW), 3))
        feat_nchw = f.permute(0, 3, 1, 2).contiguous().cpu()
    close(feat_nchw, f_ref, 'context')
    big = [dict(g) for g in (vt.mask_1_grid, vt.mask_2_grid, vt.mask_3_grid)]
    ref = O.view_transform((x,) + tuple(rig), depth.cpu(), feat_nchw, height.cpu(), fr, cfg['height_range'],
                           cfg['mask_range'], big)
    for got, want, name in zip((bev, lo, mid, hi), ref, ('bev', 'low', 'mid', 'high')):
        close(got, want, 'pooled ' + name, atol=2e-6, rtol=1e-5)
    assert vt.grid_config is vt.mask_3_grid        # the reference's leftover grid (LH:455)
