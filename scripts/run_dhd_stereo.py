"""End-to-end run of the DHD_stereo shell (DHD-L wiring: 2 temporal frames + 1 stereo reference frame, pre-process nets,
bev / voxel encoders, SFA, head) on synthetic per-frame image features at a reduced image size (64x176, C_in = 64,
stereo C = 16; BEV grid, encoder widths and head as in DHD-L.py).  Prints one JSON line.  CUDA only."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dhd_b200 import compat as C  # noqa: E402
from dhd_b200 import synth  # noqa: E402

NT = 64


def model_cfg(precision='bf16'):
    vt = dict(synth.DHD_L_VIEW_TRANSFORMER, type='MGHS_Stereo', input_size=(64, 176), in_channels=64, precision=precision)
    vt['depthnet_cfg'] = dict(use_dcn=False, aspp_mid_channels=32, stereo=True, bias=5.)
    vt['heightnet_cfg'] = dict(use_dcn=False, aspp_mid_channels=32)
    unet = lambda cin, ncls: dict(type='UNet', n_channels=cin, n_classes=ncls, precision=precision)
    pre = lambda c: dict(type='CustomResNet', numC_input=c, num_layer=[1], num_channels=[c], stride=[1],
                         backbone_output_ids=[0], precision=precision)
    return dict(
        type='DHD_stereo', align_after_view_transfromation=False, num_adj=1, img_view_transformer=vt,
        img_bev_encoder_backbone=dict(type='CustomResNet', numC_input=NT * 2, num_channels=[NT * 2, NT * 4, NT * 8],
                                      precision=precision),
        img_bev_encoder_neck=dict(type='FPN_LSS', in_channels=NT * 8 + NT * 2, out_channels=256, precision=precision),
        pre_process=pre(NT), pre_process_net_3d=pre(NT * 16),
        img_voxel_encoder0_backbone=unet(NT * 4 * 2, 64), img_voxel_encoder0_neck=dict(type='Identity'),
        img_voxel_encoder1_backbone=unet(NT * 4 * 2, 128), img_voxel_encoder1_neck=dict(type='Identity'),
        img_voxel_encoder2_backbone=unet(NT * 8 * 2, 64), img_voxel_encoder2_neck=dict(type='Identity'),
        mix=dict(type='SFA', in_channels=512, out_channels=256, precision=precision),
        occ_head=dict(type='predictor', in_dim=256, out_dim=256, Dz=16, use_mask=True, num_classes=18, use_predicter=True,
                      class_balance=False, loss_occ=None, precision=precision))


def run():
    import projects.mmdet3d_plugin  # noqa: F401
    t0 = time.time()
    torch.manual_seed(0)
    model = C.build_model(model_cfg()).eval().cuda()
    B, N = 1, 6
    rig = synth.synthetic_rig(B, N, (64, 176), seed=2)
    s2e, e2g, K, pr, pt, bda = [t.cuda() for t in rig]
    k2s = synth.synthetic_k2s_sensor(s2e)
    nf = model.num_frame                                               # key, previous, stereo reference
    feat = lambda: torch.randn(B, N, 64, 4, 11, device='cuda')
    k = torch.ones(1, 1, 3, 3, device='cuda') / 9.0
    sfeat = lambda: torch.nn.functional.conv2d(torch.randn(B * N * 16, 1, 16, 44, device='cuda'), k, padding=1).view(B * N, 16, 16, 44)
    feats = [feat(), feat(), None]
    stereo = [sfeat() for _ in range(nf)]
    per_frame = lambda t: [t] * nf
    with torch.no_grad():
        occ, depth, height = model.forward_hot_path(feats, stereo, per_frame(s2e), per_frame(e2g), per_frame(K),
                                                    per_frame(pr), per_frame(pt), bda, per_frame(k2s))
        cls = model.simple_test_occ(occ)
    torch.cuda.synchronize()
    res = {'what': 'DHD_stereo.forward_hot_path (DHD-L wiring, reduced image size)', 'occ': list(occ.shape),
           'depth': list(depth.shape), 'height': list(height.shape), 'finite': bool(torch.isfinite(occ).all()),
           'occ_abs_mean': float(occ.abs().mean()), 'classes_found': int(torch.as_tensor(cls[0]).unique().numel()),
           'seconds_total': time.time() - t0}
    return res


if __name__ == '__main__':
    print(json.dumps(run()))
