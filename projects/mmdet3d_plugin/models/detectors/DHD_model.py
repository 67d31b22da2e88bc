"""DHD detector shell (reference: projects/mmdet3d_plugin/models/detectors/DHD_model.py:10-241).

Wires the hot path exactly like DHD.extract_img_feat / forward_occ_train / simple_test_occ:
image features -> MGHS -> [BEV encoder + three voxel encoders] -> cat -> SFA -> predictor.
The image backbone / FPN and the BEV / voxel encoders are outside this build's scope
(SURVEY.md 8(f)); they are built from the registry when their `type` is registered (e.g. by a
real mmdet3d install) and otherwise left as None, in which case `forward_hot_path` takes the
encoder outputs as arguments.  Child names follow DM:22-29 so checkpoints map unchanged.
"""
import torch

from dhd_b200 import compat as C


def _maybe(registry, cfg):
    if cfg is None:
        return None
    typ = cfg.get('type')
    return registry.build(cfg) if (not isinstance(typ, str) or typ in getattr(registry, 'module_dict', {typ: 1})) else None


@C.DETECTORS.register_module(force=True)
class DHD(C.BaseModule):
    def __init__(self, img_view_transformer, mix=None, occ_head=None, img_backbone=None, img_neck=None,
                 img_bev_encoder_backbone=None, img_bev_encoder_neck=None,
                 img_voxel_encoder0_backbone=None, img_voxel_encoder0_neck=None,
                 img_voxel_encoder1_backbone=None, img_voxel_encoder1_neck=None,
                 img_voxel_encoder2_backbone=None, img_voxel_encoder2_neck=None,
                 upsample=False, train_cfg=None, test_cfg=None, **kwargs):
        super().__init__()
        self.img_backbone = _maybe(C.BACKBONES, img_backbone)
        self.img_neck = _maybe(C.NECKS, img_neck)
        self.img_view_transformer = C.NECKS.build(img_view_transformer)
        self.img_bev_encoder_backbone = _maybe(C.BACKBONES, img_bev_encoder_backbone)
        self.img_bev_encoder_neck = _maybe(C.NECKS, img_bev_encoder_neck)
        for i, (b, n) in enumerate(((img_voxel_encoder0_backbone, img_voxel_encoder0_neck),
                                    (img_voxel_encoder1_backbone, img_voxel_encoder1_neck),
                                    (img_voxel_encoder2_backbone, img_voxel_encoder2_neck))):
            setattr(self, 'img_voxel_encoder%d' % i, _maybe(C.BACKBONES, b))
            setattr(self, 'img_voxel_neck%d' % i, _maybe(C.NECKS, n))
        self.mix = C.NECKS.build(mix) if mix is not None else None
        self.occ_head = C.HEADS.build(occ_head) if occ_head is not None else None
        self.upsample = upsample
        self.train_cfg, self.test_cfg = train_cfg, test_cfg

    # DM:32-82: bev_encoder / voxel_encoder{0,1,2} = backbone -> neck (first output if a list)
    def _encode(self, backbone, neck, x):
        if backbone is None:
            raise NotImplementedError('this encoder is outside the hot path and is not registered here')
        x = backbone(x)
        if neck is not None:
            x = neck(x)
        return x[0] if isinstance(x, (list, tuple)) else x

    def bev_encoder(self, x):
        return self._encode(self.img_bev_encoder_backbone, self.img_bev_encoder_neck, x)

    def voxel_encoder(self, i, x):
        return self._encode(getattr(self, 'img_voxel_encoder%d' % i), getattr(self, 'img_voxel_neck%d' % i), x)

    def view_transform(self, img_feat, cams):
        """img_feat (B, N, C, fH, fW); cams = (sensor2egos, ego2globals, intrins, post_rots,
        post_trans, bda) -> MGHS outputs (bev, depth, height, low, mid, high), DM:84-103."""
        vt = self.img_view_transformer
        mlp_input = vt.get_mlp_input(*cams)
        return vt([img_feat] + list(cams) + [mlp_input])

    def forward_hot_path(self, img_feat, cams, encoded=None):
        """Image features -> occupancy logits (B, Dx, Dy, Dz, n_cls).  `encoded` = (x_2d, x_3d)
        overrides the encoders (each (B, 256, Dy, Dx)); DM:103-114, 196-198, 228-241."""
        bev, depth, height, low, mid, high = self.view_transform(img_feat, cams)
        if encoded is None:
            x_2d = self.bev_encoder(bev)
            x_3d = torch.cat([self.voxel_encoder(i, t) for i, t in enumerate((low, mid, high))], dim=1)
        else:
            x_2d, x_3d = encoded
        fused = self.mix(torch.cat([x_2d, x_3d], dim=1), return_act=True)
        return self.occ_head(fused), depth, height

    def simple_test_occ(self, occ_pred, img_metas=None):
        return self.occ_head.get_occ(occ_pred, img_metas)


@C.DETECTORS.register_module(force=True)
class DHD_stereo(DHD):
    """DHD-M / DHD-L detector shell (reference DHD_model.py:243-560 on BEVStereo4D <- BEVDet4D, bevstereo4d.py:13-19,
    bevdet4d.py:24-41): two-frame temporal fusion with the plane-sweep stereo DepthNet.  Child names follow the
    reference (`pre_process_net`, `pre_process_net_3d`, `img_voxel_encoder{0,1,2}`, `img_voxel_neck{0,1,2}`, `mix`,
    `occ_head`, ...), so checkpoints map unchanged.  The image backbone / neck are outside this build: the methods
    below take what `image_encoder(img, stereo=True)` returns -- the 1/16 image features and the 1/4 stereo feature of
    a frame -- instead of the images."""

    def __init__(self, pre_process=None, pre_process_net_3d=None, align_after_view_transfromation=False, num_adj=1,
                 with_prev=True, **kwargs):
        super().__init__(**kwargs)
        self.pre_process = pre_process is not None                     # bevdet4d.py:33-35
        if self.pre_process:
            self.pre_process_net = C.BACKBONES.build(pre_process)
            self.pre_process_net_3d = C.BACKBONES.build(pre_process_net_3d)      # DHD_model.py:265-266
        self.align_after_view_transfromation = align_after_view_transfromation
        self.num_frame = num_adj + 1
        self.with_prev = with_prev
        self.grid = None
        self.extra_ref_frames = 1                                      # bevstereo4d.py:16-18
        self.temporal_frame = self.num_frame
        self.num_frame += self.extra_ref_frames

    @staticmethod
    def _collapse_z(x):
        """(B, C, Dz, Dy, Dx) -> (B, Dz*C, Dy, Dx), channel = z*C + c (`torch.cat(x.unbind(dim=2), 1)`)."""
        return torch.cat(x.unbind(dim=2), 1)

    def prepare_bev_feat(self, x, stereo_feat, sensor2keyego, ego2global, intrin, post_rot, post_tran, bda, mlp_input,
                         feat_prev_iv, k2s_sensor, extra_ref_frame=False):
        """DHD_model.py:313-374 from the frame's image features.  x (B, N, C, fH, fW) (None for the extra reference
        frame, which only contributes its stereo feature); stereo_feat (B*N, C_stereo, 4fH, 4fW).
        -> (bev_feat_2d, bev_feat_3d, depth, height, stereo_feat)."""
        if extra_ref_frame:
            return None, None, None, None, stereo_feat
        vt = self.img_view_transformer
        metas = dict(k2s_sensor=k2s_sensor, intrins=intrin, post_rots=post_rot, post_trans=post_tran,
                     frustum=vt.cv_frustum.to(x), cv_downsample=4, downsample=vt.downsample,
                     grid_config=vt.grid_config, cv_feat_list=[feat_prev_iv, stereo_feat])
        bev_2d, bev_3d, depth, height = vt([x, sensor2keyego, ego2global, intrin, post_rot, post_tran, bda, mlp_input],
                                           metas)
        if self.pre_process and bev_3d.dim() == 5:
            nz = bev_3d.shape[2]
            b2 = self.pre_process_net(self._collapse_z(bev_2d))[0]
            b3 = self.pre_process_net_3d(self._collapse_z(bev_3d))[0]
            bev_2d = torch.stack(torch.chunk(b2, 1, dim=1), dim=2)
            bev_3d = torch.stack(torch.chunk(b3, nz, dim=1), dim=2)
        return bev_2d, bev_3d, depth, height, stereo_feat

    def fuse_frames(self, bev_feat_2d_list, bev_feat_3d_list):
        """DHD_model.py:517-541: frames concatenated on the channel axis, z collapsed into channels, the 16 height
        planes split 4 / 4 / 8 for the three voxel encoders.  -> (x_2d, x_3d), each (B, C_out, Dy, Dx)."""
        bev_2d = self._collapse_z(torch.cat(bev_feat_2d_list, dim=1))
        bev_3d = torch.cat(bev_feat_3d_list, dim=1)
        slabs = (bev_3d[:, :, :4], bev_3d[:, :, 4:8], bev_3d[:, :, 8:])
        x_2d = self.bev_encoder(bev_2d)
        x_3d = torch.cat([self.voxel_encoder(i, self._collapse_z(s)) for i, s in enumerate(slabs)], dim=1)
        return x_2d, x_3d

    def extract_bev_feat(self, feats, stereo_feats, sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda,
                         curr2adjsensor):
        """The frame loop of DHD_stereo.extract_img_feat (DHD_model.py:406-446, 479-541) on per-frame image features.
        All lists are indexed by frame id: 0 = key frame, 1 .. temporal_frame-1 = previous frames, last = the extra
        stereo reference frame (feats[-1] may be None).  -> (x_2d, x_3d, depth_key_frame, height_key_frame)."""
        if self.align_after_view_transfromation:
            raise NotImplementedError('shift_feature (bevdet4d.py:43-134): no DHD config aligns after the view transform')
        vt = self.img_view_transformer
        list_2d, list_3d, depth_key, height_key, feat_prev_iv = [], [], None, None, None
        for fid in range(self.num_frame - 1, -1, -1):
            key_frame = fid == 0
            extra_ref_frame = fid == self.num_frame - self.extra_ref_frames
            if not (key_frame or self.with_prev):
                continue
            x = feats[fid]
            mlp_input = None if extra_ref_frame else \
                vt.get_mlp_input(sensor2keyegos[0], ego2globals[0], intrins[fid], post_rots[fid], post_trans[fid], bda)
            with torch.set_grad_enabled(key_frame and torch.is_grad_enabled()):
                b2, b3, depth, height, feat_curr_iv = self.prepare_bev_feat(
                    x, stereo_feats[fid], sensor2keyegos[fid], ego2globals[fid], intrins[fid], post_rots[fid],
                    post_trans[fid], bda, mlp_input, feat_prev_iv, curr2adjsensor[fid], extra_ref_frame)
            if key_frame:
                depth_key, height_key = depth, height
            if not extra_ref_frame:
                list_2d.append(b2)
                list_3d.append(b3)
            if not key_frame:
                feat_prev_iv = feat_curr_iv
        if not self.with_prev:                                          # zeros in place of the previous frames (479-501)
            n_prev = self.num_frame - self.extra_ref_frames - 1
            pad = lambda t: torch.zeros([t.shape[0], t.shape[1] * n_prev] + list(t.shape[2:]), dtype=t.dtype, device=t.device)
            list_2d, list_3d = [pad(list_2d[0]), list_2d[0]], [pad(list_3d[0]), list_3d[0]]
        x_2d, x_3d = self.fuse_frames(list_2d, list_3d)
        return x_2d, x_3d, depth_key, height_key

    def forward_hot_path(self, feats, stereo_feats, sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda,
                         curr2adjsensor):
        """Per-frame image features -> occupancy logits (B, Dx, Dy, Dz, n_cls), key-frame depth and height."""
        x_2d, x_3d, depth, height = self.extract_bev_feat(feats, stereo_feats, sensor2keyegos, ego2globals, intrins,
                                                          post_rots, post_trans, bda, curr2adjsensor)
        fused = self.mix(torch.cat([x_2d, x_3d], dim=1), return_act=True)
        return self.occ_head(fused), depth, height
