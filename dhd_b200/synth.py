"""Synthetic workload of the benchmark: the DHD-S shapes of
projects/configs/DHD/DHD-S.py:20-105 and a seeded 6-camera surround rig with nuScenes-like intrinsics and
the resize + crop augmentation of the reference's data pipeline (datasets/pipelines/loading.py:55-94).
Product-side generator: bench.py's own arm and the pipeline use this one (the CPU oracle keeps an
independent copy; tests/test_oracle.py checks the two agree)."""
import math

import torch

BEV_GRID = {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [-1, 5.4, 6.4]}     # lss_heightmap.py:425-431

DHD_S = dict(
    input_size=(256, 704), downsample=16, depth=[1.0, 45.0, 1.0], C=64, C_in=256,
    height_range=[round(-1.0 + 0.1 * i, 1) for i in range(65)],
    mask_range=[-1.0, 0.6, 2.2, 5.4],
    mask_grids=[
        {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [-1, 0.6, 0.4]},
        {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [0.6, 2.2, 0.4]},
        {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [2.2, 5.4, 0.4]},
    ],
    bev_grid=BEV_GRID, ncams=6,
)


def synthetic_rig(B, ncams=6, input_size=(256, 704), src_size=(900, 1600), seed=0, flip_bda=False):
    """(sensor2ego (B,N,4,4), ego2global, cam2img (B,N,3,3), post_rot, post_tran (B,N,3), bda (B,3,3))."""
    g = torch.Generator().manual_seed(seed)
    yaws = [55.0, 0.0, -55.0, 110.0, 180.0, -110.0][:ncams]
    base = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    s2e = torch.zeros(B, ncams, 4, 4)
    for n, yaw in enumerate(yaws):
        a = math.radians(yaw)
        rz = torch.tensor([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]])
        s2e[:, n, :3, :3] = rz @ base
        s2e[:, n, :3, 3] = torch.tensor([1.5 * math.cos(a), 1.5 * math.sin(a), 1.5])
        s2e[:, n, 3, 3] = 1.0
    K = torch.tensor([[1266.0, 0.0, 816.0], [0.0, 1266.0, 491.0], [0.0, 0.0, 1.0]]).expand(B, ncams, 3, 3).clone()
    scale = input_size[1] / src_size[1]
    s = scale + 0.01 * torch.rand(B, ncams, generator=g)
    pr = torch.zeros(B, ncams, 3, 3)
    pr[..., 0, 0] = s
    pr[..., 1, 1] = s
    pr[..., 2, 2] = 1.0
    pt = torch.zeros(B, ncams, 3)
    pt[..., 1] = -(src_size[0] * scale - input_size[0])
    bda = torch.eye(3).expand(B, 3, 3).clone()
    if flip_bda:
        bda[1::2, 0, 0] = -1.0
        bda[1::2, 1, 1] = -1.0
    e2g = torch.eye(4).expand(B, ncams, 4, 4).clone()
    return s2e, e2g, K, pr, pt, bda


# ---------------------------------------------------------------------------------------------- DHD-L (cfg-5)
DHD_L_VIEW_TRANSFORMER = dict(      # kwargs of model.img_view_transformer in projects/configs/DHD/DHD-L.py:75-119
    grid_config={'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [-1, 5.4, 6.4], 'depth': [1.0, 45.0, 0.5]},
    input_size=(512, 1408),
    height_range=[round(-1.0 + 0.1 * i, 1) for i in range(65)], height_interval=0.1,
    mask_range=[-1.0, 0.6, 2.2, 5.4],
    mask_1_grid={'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [-1, 0.6, 0.4], 'depth': [1.0, 45.0, 0.5]},
    mask_2_grid={'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [0.6, 2.2, 0.4], 'depth': [1.0, 45.0, 0.5]},
    mask_3_grid={'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [2.2, 5.4, 0.4], 'depth': [1.0, 45.0, 0.5]},
    in_channels=512, out_channels=64, sid=False, collapse_z=False, loss_height_weight=0.1, loss_depth_weight=0.05,
    depthnet_cfg=dict(use_dcn=False, aspp_mid_channels=96, stereo=True, bias=5.),
    heightnet_cfg=dict(use_dcn=False, aspp_mid_channels=96),
    downsample=16)
DHD_L_STEREO_CHANNELS = 128         # Swin-B stage-0 width (DHD-L.py:44-63, return_stereo_feat=True)


def synthetic_k2s_sensor(sensor2ego, forward_m=0.8, yaw_deg=1.5):
    """Current-camera -> previous-frame-camera transforms (B, N, 4, 4) of a vehicle that drove `forward_m` metres
    while yawing `yaw_deg`: inv(sensor2ego) @ key_ego->previous_ego @ sensor2ego (the k2s_sensor the reference derives
    from the nuScenes poses, detectors/bevstereo4d.py)."""
    a = math.radians(yaw_deg)
    ego = torch.eye(4)
    ego[:3, :3] = torch.tensor([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]])
    ego[:3, 3] = torch.tensor([forward_m, 0.03, 0.0])
    ego = ego.to(sensor2ego)
    return torch.linalg.inv(sensor2ego) @ ego @ sensor2ego


def dhdl_view_transformer(precision, B):
    """The plugin's MGHS_Stereo built with the DHD-L.py kwargs + synthetic inputs of BASELINE configs[4]: (module,
    forward `input` list, stereo_metas).  CUDA only."""
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.necks.lss_heightmap import MGHS_Stereo
    kw = dict(DHD_L_VIEW_TRANSFORMER)
    torch.manual_seed(0)
    vt = MGHS_Stereo(precision=precision, **kw).eval().cuda()
    H, W = kw['input_size']
    rig = synthetic_rig(B, 6, kw['input_size'], seed=1)
    s2e, e2g, K, pr, pt, bda = [t.cuda() for t in rig]
    x = torch.randn(B, 6, kw['in_channels'], H // 16, W // 16, device='cuda')
    mlp = vt.get_mlp_input(s2e, e2g, K, pr, pt, bda)
    k = torch.ones(1, 1, 5, 5, device='cuda') / 25.0
    C = DHD_L_STEREO_CHANNELS
    feat = lambda: torch.nn.functional.conv2d(torch.randn(B * 6 * C, 1, H // 4, W // 4, device='cuda'), k,
                                              padding=2).view(B * 6, C, H // 4, W // 4)
    metas = dict(k2s_sensor=synthetic_k2s_sensor(s2e), intrins=K, post_rots=pr, post_trans=pt,
                 frustum=vt.cv_frustum.cuda(), cv_downsample=4, downsample=vt.downsample, grid_config=vt.grid_config,
                 cv_feat_list=[feat(), feat()])
    return vt, [x, s2e, e2g, K, pr, pt, bda, mlp], metas


# ---------------------------------------------------------------------------------------------- cfg-4 ("DHD-B") / cfg-5 model dicts
DHD_B = dict(DHD_S, input_size=(384, 1056))       # BASELINE configs[3]: the DHD-S topology at 384x1056 (24x66 features)


def dhd_l_model_cfg(precision='bf16'):
    """`model` of projects/configs/DHD/DHD-L.py:41-190 without the image backbone / neck (Swin-B + FPN_LSS, outside the hot
    path: the detector takes the per-frame image features instead of images)."""
    c = 64
    vt = dict(DHD_L_VIEW_TRANSFORMER, type='MGHS_Stereo', precision=precision)
    enc = lambda n_in, n_out: dict(type='UNet', n_channels=n_in, n_classes=n_out, precision=precision)
    return dict(
        type='DHD_stereo', align_after_view_transfromation=False, num_adj=1,
        img_view_transformer=vt,
        img_bev_encoder_backbone=dict(type='CustomResNet', numC_input=c * 2, num_channels=[c * 2, c * 4, c * 8], precision=precision),
        img_bev_encoder_neck=dict(type='FPN_LSS', in_channels=c * 8 + c * 2, out_channels=256, precision=precision),
        pre_process=dict(type='CustomResNet', numC_input=c, num_layer=[1], num_channels=[c], stride=[1], backbone_output_ids=[0],
                         precision=precision),
        pre_process_net_3d=dict(type='CustomResNet', numC_input=c * 16, num_layer=[1], num_channels=[c * 16], stride=[1],
                                backbone_output_ids=[0], precision=precision),
        img_voxel_encoder0_backbone=enc(c * 8, 64), img_voxel_encoder0_neck=dict(type='Identity'),
        img_voxel_encoder1_backbone=enc(c * 8, 128), img_voxel_encoder1_neck=dict(type='Identity'),
        img_voxel_encoder2_backbone=enc(c * 16, 64), img_voxel_encoder2_neck=dict(type='Identity'),
        mix=dict(type='SFA', in_channels=512, out_channels=256, precision=precision),
        occ_head=dict(type='predictor', in_dim=256, out_dim=256, Dz=16, use_mask=True, num_classes=18, use_predicter=True,
                      class_balance=True, weight_ce=10.0, precision=precision,
                      loss_occ=dict(type='CrossEntropyLoss', use_sigmoid=False, ignore_index=255, loss_weight=1.0)))


def dhd_s_model_cfg(precision='bf16', images=True, input_size=(256, 704), depth=50):
    """`model` of projects/configs/DHD/DHD-S.py:41-155: with images=True INCLUDING the image backbone / neck entries
    (DHD-S.py:44-62: mmdet ResNet-50 trained with batch-statistics BatchNorm, CustomFPN) -- the detector then takes the
    camera images, as under the reference's runner; depth=101 + input_size=(384, 1056) is BASELINE configs[3] ("DHD-B")."""
    c = 64
    grid = {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [-1, 5.4, 6.4], 'depth': [1.0, 45.0, 1.0]}
    mg = lambda z: {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': z, 'depth': [1.0, 45.0, 0.5]}
    enc = lambda n_in, n_out: dict(type='UNet', n_channels=n_in, n_classes=n_out, precision=precision)
    cfg = dict(
        type='DHD',
        img_view_transformer=dict(type='MGHS', grid_config=grid, input_size=tuple(input_size),
                                  height_range=[round(-1.0 + 0.1 * i, 1) for i in range(65)], height_interval=0.1,
                                  mask_1_grid=mg([-1, 0.6, 0.4]), mask_2_grid=mg([0.6, 2.2, 0.4]), mask_3_grid=mg([2.2, 5.4, 0.4]),
                                  mask_range=[-1.0, 0.6, 2.2, 5.4], loss_height_weight=0.1, in_channels=256, out_channels=c,
                                  sid=False, collapse_z=True, downsample=16, precision=precision),
        img_bev_encoder_backbone=dict(type='CustomResNet', numC_input=c, num_channels=[c * 2, c * 4, c * 8], precision=precision),
        img_bev_encoder_neck=dict(type='FPN_LSS', in_channels=c * 8 + c * 2, out_channels=256, precision=precision),
        img_voxel_encoder0_backbone=enc(c * 4, 64), img_voxel_encoder0_neck=dict(type='Identity'),
        img_voxel_encoder1_backbone=enc(c * 4, 128), img_voxel_encoder1_neck=dict(type='Identity'),
        img_voxel_encoder2_backbone=enc(c * 8, 64), img_voxel_encoder2_neck=dict(type='Identity'),
        mix=dict(type='SFA', in_channels=512, out_channels=256, precision=precision),
        occ_head=dict(type='predictor', in_dim=256, out_dim=256, Dz=16, use_mask=True, num_classes=18, use_predicter=True,
                      class_balance=True, precision=precision,
                      loss_occ=dict(type='CrossEntropyLoss', use_sigmoid=False, ignore_index=255, loss_weight=1.0)))
    if images:
        cfg['img_backbone'] = dict(type='ResNet', depth=depth, num_stages=4, out_indices=(2, 3), frozen_stages=-1,
                                   norm_cfg=dict(type='BN', requires_grad=True), norm_eval=False, with_cp=True, style='pytorch',
                                   pretrained='torchvision://resnet%d' % depth, precision=precision)
        cfg['img_neck'] = dict(type='CustomFPN', in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0,
                               out_ids=[0], precision=precision)
    return cfg
