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
