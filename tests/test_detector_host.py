"""CPU tests of the detectors' host-side layout logic (no GPU): the activation fast path of DHD_stereo.simple_test
re-orders channels with plain tensor ops that must reproduce the reference's cat / unbind sequence
(detectors/DHD_model.py:313-374, 517-541 of the reference)."""
import torch


def test_frame_fusion_on_channels_last_rows_matches_the_reference_cat_unbind_sequence():
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.detectors.DHD_model import DHD_stereo
    g = torch.Generator().manual_seed(0)
    B, C, nz, H, W, F = 2, 8, 16, 5, 7, 2
    b2 = [torch.randn(B, C, 1, H, W, generator=g) for _ in range(F)]          # per frame, as the reference holds them
    b3 = [torch.randn(B, C, nz, H, W, generator=g) for _ in range(F)]
    collapse = DHD_stereo._collapse_z                                         # torch.cat(x.unbind(dim=2), 1)
    # reference: frames concatenated on C, z collapsed into channels, the planes split 4 / 4 / 8
    want_2d = collapse(torch.cat(b2, dim=1))
    cat3 = torch.cat(b3, dim=1)
    want_slabs = [collapse(s) for s in (cat3[:, :, :4], cat3[:, :, 4:8], cat3[:, :, 8:])]
    # fast path: every frame arrives as channels-last rows with channel = z*C + c (what the pool kernel writes)
    rows_2d = [collapse(t).permute(0, 2, 3, 1).contiguous() for t in b2]
    rows_3d = [collapse(t).permute(0, 2, 3, 1).contiguous() for t in b3]
    assert torch.equal(DHD_stereo._frames_to_bev_rows(rows_2d), want_2d.permute(0, 2, 3, 1))
    for (z0, z1), want in zip(((0, 4), (4, 8), (8, nz)), want_slabs):
        got = DHD_stereo._frames_to_slab_rows(rows_3d, C, z0, z1)
        assert got.shape == (B, H, W, (z1 - z0) * F * C)
        assert torch.equal(got, want.permute(0, 2, 3, 1))


def test_act_path_is_selected_only_for_the_bf16_speed_mode():
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200 import compat as C
    from dhd_b200 import synth
    m = C.DETECTORS.build(synth.dhd_s_model_cfg('bf16', images=False))
    assert m.eval()._dhd_act_path_ok() and not m.train()._dhd_act_path_ok()
    m.eval().act_path = False
    assert not m._dhd_act_path_ok()
    assert not C.DETECTORS.build(synth.dhd_s_model_cfg('fp32', images=False)).eval()._dhd_act_path_ok()
    cfg = synth.dhd_l_model_cfg('bf16')
    cfg['img_view_transformer'] = dict(cfg['img_view_transformer'], input_size=(128, 352))
    s = C.DETECTORS.build(cfg)
    assert s.eval()._act_path_ok() and not s.train()._act_path_ok()
