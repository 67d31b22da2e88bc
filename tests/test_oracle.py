"""CPU tests: pin the oracle (a) to the reference's only known-answer vector, (b) to the
fixtures generated from the real reference code, (c) to the real reference itself where
/root/reference exists (this container)."""
import numpy as np
import pytest
import torch

from oracle import mghs_oracle as O
from oracle import ref_loader
from tests import helpers as H


def test_kat_bev_pool_v2_reference_vector():
    """The reference's own test_bev_pool_v2 (ops/bev_pool_v2/bev_pool.py:163-194)."""
    depth = torch.tensor([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9]).view(1, 1, 2, 2, 2).requires_grad_()
    feat = torch.ones(1, 1, 2, 2, 2, requires_grad=True)
    rd = torch.tensor([0, 4, 1, 6], dtype=torch.int32)
    rf = torch.tensor([0, 0, 1, 2], dtype=torch.int32)
    rb = torch.tensor([0, 0, 1, 1], dtype=torch.int32)
    st = torch.tensor([0, 2], dtype=torch.int32)
    ln = torch.tensor([2, 2], dtype=torch.int32)
    out = O.bev_pool_v2(depth, feat, rd, rf, rb, (1, 1, 2, 2, 2), st, ln)
    loss = out.sum()
    loss.backward()
    assert loss.item() == pytest.approx(4.4, abs=1e-6)
    assert torch.allclose(depth.grad.flatten(), torch.tensor([2., 2., 0., 0., 2., 0., 2., 0.]))
    assert torch.allclose(feat.grad.flatten(), torch.tensor([1.0, 1.0, 0.4, 0.4, 0.8, 0.8, 0., 0.]))


def test_c_and_index_add_restatements_agree():
    cfg, B, inputs, depth, feat, height, gold = H.load_case('cfg1_b1')
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample'])
    coor = O.ego_coor(fr, inputs[1], inputs[3], inputs[4], inputs[5], inputs[6])
    lower, interval, size = O.grid_infos(**cfg['bev_grid'])
    rb, rd, rf, st, ln = O.prepare_v2(coor, lower, interval, size)
    N, D = cfg['ncams'], depth.shape[1]
    fH, fW = depth.shape[-2:]
    d5 = depth.view(B, N, D, fH, fW)
    f5 = feat.view(B, N, cfg['C'], fH, fW).permute(0, 1, 3, 4, 2).contiguous()
    shape = (B, int(size[2]), int(size[1]), int(size[0]), cfg['C'])
    a = O.bev_pool_v2(d5, f5, rd, rf, rb, shape, st, ln)
    b = O.bev_pool_v2_index_add(d5, f5, rd, rf, rb, shape)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('name', ['cfg1_b1', 'mini_mghs_b2', 'dhds_b1', 'dhds_b2_flip'])
def test_oracle_indices_match_reference_fixture(name):
    cfg, B, inputs, depth, feat, height, gold = H.load_case(name)
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample'])
    coor = O.ego_coor(fr, inputs[1], inputs[3], inputs[4], inputs[5], inputs[6])
    assert H.sha(coor) == str(gold['coor_sha'])          # geometry bit-exact
    for p, g in enumerate(H.grids_of(cfg)):
        r = H.oracle_ranks(coor, g)
        if 'ranks_%d' % p in gold:
            assert np.array_equal(r.numpy(), gold['ranks_%d' % p])
        else:
            assert H.sha(r) == str(gold['ranks_sha_%d' % p])
        lower, interval, size = O.grid_infos(g['x'], g['y'], g['z'])
        rb, rd, rf, st, ln = O.prepare_v2(coor, lower, interval, size)
        assert st.numel() == int(gold['n_intervals_%d' % p])
        assert rb.numel() == int(gold['n_kept_%d' % p])
    if height is not None:
        mid, _ = O.height_masks(height, cfg['height_range'], cfg['mask_range'])
        assert H.sha(mid) == str(gold['mask_id_sha'])


@pytest.mark.parametrize('name', ['cfg1_b1', 'mini_mghs_b2'])
def test_oracle_pool_matches_reference_fixture(name):
    cfg, B, inputs, depth, feat, height, gold = H.load_case(name)
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample'])
    if height is None:
        coor = O.ego_coor(fr, inputs[1], inputs[3], inputs[4], inputs[5], inputs[6])
        N, D = cfg['ncams'], depth.shape[1]
        fH, fW = depth.shape[-2:]
        outs = (O.pool_one_pass(coor, depth.view(B, N, D, fH, fW),
                                feat.view(B, N, cfg['C'], fH, fW), cfg['bev_grid']),)
    else:
        outs = O.view_transform(inputs, depth, feat, height, fr, cfg['height_range'],
                                cfg['mask_range'], cfg['mask_grids'], bev_grid=cfg['bev_grid'])
    for p, o in enumerate(outs):
        ref = torch.from_numpy(gold['out_%d' % p])
        assert o.shape == ref.shape
        assert torch.equal(o != 0, ref != 0)
        # summation order inside an interval is undefined in the reference (unstable argsort)
        assert torch.allclose(o, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_oracle_vs_live_reference_view_transform():
    import warnings
    from oracle.make_golden import ref_geometry_object
    ns = ref_loader.load_reference()
    cfg = O.MINI
    inputs, depth, feat, height = O.synthetic_inputs(cfg, 1, seed=11, flip_bda=False)
    m = ref_geometry_object(ns, cfg)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = m.view_transform(list(inputs), depth, feat, height)
        coor = m.get_ego_coor(*inputs[1:7])
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample'])
    assert torch.equal(fr, m.frustum)
    assert torch.equal(coor, O.ego_coor(fr, inputs[1], inputs[3], inputs[4], inputs[5], inputs[6]))
    ours = O.view_transform(inputs, depth, feat, height, fr, cfg['height_range'],
                            cfg['mask_range'], cfg['mask_grids'], bev_grid=cfg['bev_grid'])
    for a, b in zip((ref[0], ref[3], ref[4], ref[5]), ours):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    # reference prepare: indices bit-exact, (rank_depth, rank_feat) equal as per-interval sets
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m.create_grid_infos(**cfg['mask_grids'][1])
        r = m.voxel_pooling_prepare_v2(coor)
    lo, iv, sz = O.grid_infos(**cfg['mask_grids'][1])
    o = O.prepare_v2(coor, lo, iv, sz)
    assert torch.equal(r[0], o[0]) and torch.equal(r[3], o[3]) and torch.equal(r[4], o[4])
    key_r = r[0].long() * (1 << 32) + r[1].long()
    key_o = o[0].long() * (1 << 32) + o[1].long()
    assert torch.equal(key_r.sort().values, key_o.sort().values)


def test_product_synthetic_workload_matches_the_oracle_copy():
    """bench.py's own arm builds its inputs with dhd_b200.synth (it must not import oracle/); the oracle
    keeps an independent copy of the same rig -- they have to describe the same workload."""
    from dhd_b200 import synth
    assert synth.DHD_S == O.DHD_S
    for seed, flip in ((0, False), (103, True)):
        a = synth.synthetic_rig(3, 6, (256, 704), seed=seed, flip_bda=flip)
        b = O.synthetic_rig(3, 6, (256, 704), seed=seed, flip_bda=flip)
        for x, y in zip(a, b):
            assert torch.equal(x, y)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_loss_oracle_matches_reference_losses():
    """oracle/loss_oracle.py (vectorised) against the unmodified semkitti_loss functions: values and gradients."""
    from oracle import loss_oracle as LO
    sk = ref_loader.load_reference().semkitti_loss
    g = torch.Generator().manual_seed(4)
    n = 4000
    labels = torch.randint(0, 18, (n,), generator=g)
    labels[:50] = 255
    labels[labels == 5] = 6                         # one class absent from the target
    mask = torch.rand(n, generator=g) < 0.6
    for scale in (1.0, 6.0):
        logits = (torch.randn(n, 18, generator=g) * scale)
        for ref_fn, our_fn, kw in ((sk.sem_scal_loss_with_mask, LO.sem_scal_loss_with_mask, {}),
                                   (sk.geo_scal_loss_with_mask, LO.geo_scal_loss_with_mask, dict(non_empty_idx=17))):
            a = logits.clone().requires_grad_()
            b = logits.clone().requires_grad_()
            la = ref_fn(a, labels, mask.int(), **kw)
            lb = our_fn(b, labels, mask, **kw)
            la.backward()
            lb.backward()
            assert torch.allclose(la, lb, rtol=1e-5, atol=1e-6), (float(la), float(lb))
            assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-9)
