"""DHD detectors (reference: projects/mmdet3d_plugin/models/detectors/DHD_model.py:10-241 `DHD`, 243-560 `DHD_stereo`).

Same registry names, constructor kwargs, child names (DM:22-29: checkpoints map unchanged) and public methods as the
reference -- extract_img_feat / extract_feat / forward_train / forward_occ_train / simple_test / simple_test_occ /
forward_test / forward(return_loss) / train_step -- wired over the B200 modules: image features -> MGHS -> BEV encoder +
three voxel encoders -> cat -> SFA -> predictor.  Every module is differentiable under autograd (dhd_b200.autograd), so
`forward_train(...)` returns the reference's loss dict and `.backward()` runs the hand-written backward kernels.
The image backbone / FPN (mmdet ResNet / Swin + CustomFPN) are outside the hot path: they are built
from the registry when their `type` is registered (a real mmdet3d install) and are a `MissingModule` placeholder
otherwise; `image_encoder` then expects the (B, N, C, fH, fW) image features in the `imgs` slot of `img_inputs`.
"""
from collections import OrderedDict

import torch

from dhd_b200 import compat as C


class MissingModule(torch.nn.Module):
    """Placeholder for a config entry whose `type` is in no registry of this environment (the image backbone / FPN of
    the reference come from mmdet, which is not installed here and outside the hot path): the detector still builds
    with the reference's child names, and calling the placeholder says exactly what is missing."""

    def __init__(self, typ, registry):
        super().__init__()
        self.missing_type, self.registry = str(typ), registry

    def forward(self, *a, **k):
        raise NotImplementedError('%s is not registered in the %s registry of this environment (image backbone / FPN '
                                  'are outside the dhd_b200 hot path): install mmdet3d, register the class, or hand '
                                  'DHD the (B, N, C, fH, fW) image features in place of the images' %
                                  (self.missing_type, self.registry))


def _maybe(registry, cfg):
    """Build `cfg` from `registry` (parent registries included when it is an mmcv registry); an unknown type gives a
    MissingModule placeholder that raises when it is called, never a silent None."""
    if cfg is None:
        return None
    typ = cfg.get('type')
    if isinstance(typ, str) and registry.get(typ) is None:
        return MissingModule(typ, getattr(registry, 'name', '?'))
    return registry.build(cfg)


@C.DETECTORS.register_module(force=True)
class DHD(C.BaseModule):
    """reference DHD_model.py:10-241 (on BEVDet <- CenterPoint <- MVXTwoStageDetector <- Base3DDetector): the same
    public methods with the same signatures and loss-dict keys, so the reference's runner (`model.train_step` /
    `model(return_loss=..., **data)`, tools/train.py:276, tools/test.py:267) drives the B200 path."""

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

    @property
    def with_img_neck(self):
        return self.img_neck is not None

    # ------------------------------------------------------------------ BEVDet pieces (bevdet.py:21-78)
    def image_encoder(self, img, stereo=False):
        """bevdet.py:21-44: (B, N, 3, H, W) images -> (B, N, C, fH, fW) features [+ the 1/4 stereo feature].
        The image backbone / neck are outside this build: when they are not registered (MissingModule) the argument
        must already be the (B, N, C_in, fH, fW) feature map the view transformer reads -- the injection point for an
        external backbone."""
        B, N, C, imH, imW = img.shape
        vt = self.img_view_transformer
        fH, fW = vt.input_size[0] // vt.downsample, vt.input_size[1] // vt.downsample
        if C == vt.in_channels and (imH, imW) == (fH, fW) and not stereo:
            return img, None                              # already the feature map the view transformer reads
        if self.img_backbone is None or isinstance(self.img_backbone, MissingModule):
            if C == vt.in_channels and (imH, imW) == (fH, fW):
                return img, None
            if self.img_backbone is None:
                raise NotImplementedError('DHD was built without img_backbone: pass (B, N, %d, %d, %d) image features' %
                                          (vt.in_channels, fH, fW))
            return self.img_backbone(img), None            # raises with the explanation
        x = self.img_backbone(img.view(B * N, C, imH, imW))
        stereo_feat = None
        if stereo:
            stereo_feat, x = x[0], x[1:]
        if self.with_img_neck:
            x = self.img_neck(x)
            if type(x) in [list, tuple]:
                x = x[0]
        return x.view(B, N, x.shape[1], x.shape[2], x.shape[3]), stereo_feat

    def prepare_inputs(self, inputs):
        """bevdet.py:60-78: sensor -> key-ego transforms (fp64 inverse + products, as the reference; `linalg.inv_ex` is
        `torch.inverse` without the singularity check's device -> host synchronisation: same bits, and the step stays
        capturable in a CUDA graph)."""
        assert len(inputs) == 7
        B, N = inputs[0].shape[:2]
        imgs, sensor2egos, ego2globals, intrins, post_rots, post_trans, bda = inputs
        sensor2egos = sensor2egos.view(B, N, 4, 4)
        ego2globals = ego2globals.view(B, N, 4, 4)
        keyego2global = ego2globals[:, 0, ...].unsqueeze(1)
        global2keyego = torch.linalg.inv_ex(keyego2global.double()).inverse
        sensor2keyegos = (global2keyego @ ego2globals.double() @ sensor2egos.double()).float()
        return [imgs, sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda]

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

    def voxel_encoder0(self, x):
        return self.voxel_encoder(0, x)

    def voxel_encoder1(self, x):
        return self.voxel_encoder(1, x)

    def voxel_encoder2(self, x):
        return self.voxel_encoder(2, x)

    def view_transform(self, img_feat, cams):
        """img_feat (B, N, C, fH, fW); cams = (sensor2egos, ego2globals, intrins, post_rots,
        post_trans, bda) -> MGHS outputs (bev, depth, height, low, mid, high), DM:84-103."""
        vt = self.img_view_transformer
        mlp_input = vt.get_mlp_input(*cams)
        return vt([img_feat] + list(cams) + [mlp_input])

    # ------------------------------------------------------------------ reference detector API (DM:84-241)
    def extract_img_feat(self, img_inputs, img_metas=None, **kwargs):
        """DM:84-114: img_inputs = (imgs, sensor2egos, ego2globals, intrins, post_rots, post_trans, bda) ->
        (x_2d (B, 256, Dy, Dx), x_3d (B, 256, Dy, Dx), depth (B*N, D, fH, fW), height (B*N, H, fH, fW))."""
        imgs, sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda = self.prepare_inputs(list(img_inputs))
        x, _ = self.image_encoder(imgs)
        x_2d, depth, height, mask_1, mask_2, mask_3 = self.view_transform(
            x, (sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda))
        x_2d = self.bev_encoder(x_2d)
        x_3d = torch.cat((self.voxel_encoder0(mask_1), self.voxel_encoder1(mask_2), self.voxel_encoder2(mask_3)), dim=1)
        return x_2d, x_3d, depth, height

    def extract_feat(self, points, img_inputs, img_metas=None, **kwargs):
        """DM:116-133 -> (x_2d, x_3d, pts_feats=None, depth, height)."""
        x_2d, x_3d, depth, height = self.extract_img_feat(img_inputs, img_metas, **kwargs)
        return x_2d, x_3d, None, depth, height

    def forward_train(self, points=None, img_metas=None, gt_bboxes_3d=None, gt_labels_3d=None, gt_labels=None,
                      gt_bboxes=None, img_inputs=None, proposals=None, gt_bboxes_ignore=None, **kwargs):
        """DM:135-186 -> dict(loss_height, loss_occ, loss_voxel_sem_scal, loss_voxel_geo_scal); kwargs carry
        voxel_semantics / mask_camera (B, Dx, Dy, Dz), gt_depth / gt_height (B, N, H_in, W_in)."""
        x_2d, x_3d, _pts, depth, height = self.extract_feat(points, img_inputs=img_inputs, img_metas=img_metas, **kwargs)
        losses = dict()
        losses['loss_height'] = self.img_view_transformer.get_height_loss(kwargs['gt_depth'], kwargs['gt_height'], height)
        losses.update(self.forward_occ_train([x_2d, x_3d], kwargs['voxel_semantics'], kwargs['mask_camera']))
        return losses

    def forward_occ_train(self, img_feats, voxel_semantics, mask_camera):
        """DM:188-205."""
        outs = self.occ_head(self.mix(torch.cat(img_feats, dim=1)))
        return self.occ_head.loss(outs, voxel_semantics, mask_camera)

    def simple_test(self, points, img_metas, img=None, rescale=False, to_host=True, **kwargs):
        """DM:207-226 -> list of (Dx, Dy, Dz) uint8 class maps.  to_host=False: the (B, Dx, Dy, Dz) uint8 DEVICE tensor
        instead (no synchronisation: what a CUDA-graph capture of the step needs, dhd_b200.detector_step).
        In the bf16 speed mode the chain pool -> encoders -> SFA -> head runs on bf16 NHWC activations end to end
        (`act_path = False` or DHD_ACT_PATH=0 selects the module-by-module tensor path of the reference's code)."""
        if type(self) is DHD and self._dhd_act_path_ok():
            with torch.no_grad():
                occ = self.occ_head.forward_occ(self.mix(self._extract_acts(img), return_act=True))
            return self.occ_head.get_occ(occ, img_metas) if to_host else occ
        x_2d, x_3d, _, _, _ = self.extract_feat(points, img_inputs=img, img_metas=img_metas, **kwargs)
        return self.simple_test_occ([x_2d, x_3d], img_metas, to_host=to_host)

    def _dhd_act_path_ok(self):
        import os
        if os.environ.get('DHD_ACT_PATH', '1') == '0' or not getattr(self, 'act_path', True) or self.training:
            return False
        vt = self.img_view_transformer
        mods = [vt, self.img_bev_encoder_backbone, self.img_bev_encoder_neck, self.mix, self.occ_head] + \
            [getattr(self, 'img_voxel_encoder%d' % i, None) for i in range(3)]
        if any(m is None or getattr(m, 'precision', None) != 'bf16' for m in mods) or not getattr(vt, 'collapse_z', False):
            return False
        if type(vt).__name__ != 'MGHS' or any(type(getattr(self, 'img_voxel_neck%d' % i, None)).__name__ != 'Identity'
                                               for i in range(3)):
            return False
        widths = sum(getattr(self, 'img_voxel_encoder%d' % i).n_classes for i in range(3))
        return widths == 256 and getattr(self.mix, 'mix_channels', None) == 512

    def _extract_acts(self, img_inputs):
        """extract_img_feat (DM:84-114) on activations: -> the SFA's 512-channel bf16 NHWC input, every encoder writing
        its channel slice in place (the order of DM:103-114's cat: BEV encoder | voxel encoders 0, 1, 2)."""
        from dhd_b200 import dense as D
        imgs, sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda = self.prepare_inputs(list(img_inputs))
        x, _ = self.image_encoder(imgs)
        vt = self.img_view_transformer
        cams = (sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda)
        bev, _, _, low, mid, high = vt([x] + list(cams) + [vt.get_mlp_input(*cams)], return_act=True)
        enc = D.Act.empty(bev.N, bev.H, bev.W, 512, 1, bev.data.device)
        feats = self.img_bev_encoder_backbone(bev, return_act=True)
        self.img_bev_encoder_neck(feats, return_act=True, out=enc.slice(0, 256))
        lo = 256
        for i, a in enumerate((low, mid, high)):
            net = getattr(self, 'img_voxel_encoder%d' % i)
            net(a, return_act=True, out=enc.slice(lo, lo + net.n_classes))
            lo += net.n_classes
        return enc

    def simple_test_occ(self, img_feats, img_metas=None, to_host=True):
        """DM:228-241: cat -> mix -> occ_head -> get_occ.  `img_feats` may also be the logits tensor itself (the
        round-1 calling convention).  The head's fused inference tail produces the class map directly."""
        if isinstance(img_feats, torch.Tensor):
            return self.occ_head.get_occ(img_feats, img_metas)
        fused = self.mix(torch.cat(list(img_feats), dim=1), return_act=True)
        occ = self.occ_head.forward_occ(fused)
        return self.occ_head.get_occ(occ, img_metas) if to_host else occ

    def forward_test(self, points=None, img_inputs=None, img_metas=None, **kwargs):
        """bevdet.py:168-204: one test-time augmentation only (aug_test asserts False in the reference too)."""
        for var, name in [(img_inputs, 'img_inputs'), (img_metas, 'img_metas')]:
            if not isinstance(var, list):
                raise TypeError('{} must be a list, but got {}'.format(name, type(var)))
        if len(img_inputs) != len(img_metas):
            raise ValueError('num of augmentations ({}) != num of image meta ({})'.format(len(img_inputs), len(img_metas)))
        if isinstance(img_inputs[0][0], list):
            raise AssertionError('aug_test is not implemented (reference: bevdet.py:206-208)')
        points = [points] if points is None else points
        return self.simple_test(points[0], img_metas[0], img_inputs[0], **kwargs)

    def forward(self, return_loss=True, **kwargs):
        """mmdet3d Base3DDetector.forward: forward_train with the losses, forward_test otherwise."""
        return self.forward_train(**kwargs) if return_loss else self.forward_test(**kwargs)

    @staticmethod
    def _parse_losses(losses):
        """mmdet BaseDetector._parse_losses: total of every entry whose key contains 'loss' + scalar log values."""
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError('%s is not a tensor or list of tensors' % name)
        loss = sum(v for k, v in log_vars.items() if 'loss' in k)
        log_vars['loss'] = loss
        return loss, OrderedDict((k, float(v.item())) for k, v in log_vars.items())

    def train_step(self, data, optimizer=None):
        """mmdet BaseDetector.train_step: what mmcv's EpochBasedRunner calls every iteration."""
        loss, log_vars = self._parse_losses(self(**data))
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data['img_metas']) if 'img_metas' in data else
                    int(data['img_inputs'][0].shape[0]))

    def val_step(self, data, optimizer=None):
        return self.train_step(data, optimizer)

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

    def _cv_frustum(self, like):
        """The stereo frustum template on the features' device / dtype, copied once (the module keeps it as a plain CPU
        tensor like the reference; a host -> device copy per frame would also break CUDA-graph capture)."""
        key = (str(like.device), like.dtype)
        c = self.__dict__.get('_cv_frustum_cache')
        if c is None or c[0] != key:
            c = (key, self.img_view_transformer.cv_frustum.to(like))
            self.__dict__['_cv_frustum_cache'] = c
        return c[1]

    def prepare_bev_feat(self, x, stereo_feat, sensor2keyego, ego2global, intrin, post_rot, post_tran, bda, mlp_input,
                         feat_prev_iv, k2s_sensor, extra_ref_frame=False):
        """DHD_model.py:313-374 from the frame's image features.  x (B, N, C, fH, fW) (None for the extra reference
        frame, which only contributes its stereo feature); stereo_feat (B*N, C_stereo, 4fH, 4fW).
        -> (bev_feat_2d, bev_feat_3d, depth, height, stereo_feat)."""
        if extra_ref_frame:
            return None, None, None, None, stereo_feat
        vt = self.img_view_transformer
        metas = dict(k2s_sensor=k2s_sensor, intrins=intrin, post_rots=post_rot, post_trans=post_tran,
                     frustum=self._cv_frustum(x), cv_downsample=4, downsample=vt.downsample,
                     grid_config=vt.grid_config, cv_feat_list=[feat_prev_iv, stereo_feat])
        if getattr(self, '_act_path', False):
            # inference fast path: bf16 NHWC activations from the pool kernel to the SFA, no fp32 / NCDHW round trips
            bev_2d, bev_3d, depth, height = vt([x, sensor2keyego, ego2global, intrin, post_rot, post_tran, bda, mlp_input],
                                               metas, return_act=True)
            bev_2d = self.pre_process_net(bev_2d, return_act=True)[0]
            bev_3d = self.pre_process_net_3d(bev_3d, return_act=True)[0]
            return bev_2d, bev_3d, depth, height, stereo_feat
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
        if getattr(self, '_act_path', False):
            return self._fuse_frames_acts(bev_feat_2d_list, bev_feat_3d_list), None
        bev_2d = self._collapse_z(torch.cat(bev_feat_2d_list, dim=1))
        bev_3d = torch.cat(bev_feat_3d_list, dim=1)
        slabs = (bev_3d[:, :, :4], bev_3d[:, :, 4:8], bev_3d[:, :, 8:])
        x_2d = self.bev_encoder(bev_2d)
        x_3d = torch.cat([self.voxel_encoder(i, self._collapse_z(s)) for i, s in enumerate(slabs)], dim=1)
        return x_2d, x_3d

    @staticmethod
    def _frames_to_bev_rows(rows_2d):
        """Per-frame (B, H, W, C) rows -> the BEV encoder's (B, H, W, F*C) input: `collapse_z(cat(frames, dim=1))` of
        DM:517-541 for the one-plane tensors, channels-last."""
        return torch.cat(list(rows_2d), dim=-1).contiguous()

    @staticmethod
    def _frames_to_slab_rows(rows_3d, C, z0, z1):
        """Per-frame (B, H, W, nz*C) rows (channel = z*C + c) -> the voxel encoder input of the z slab [z0, z1):
        `collapse_z(cat(frames, dim=1)[:, :, z0:z1])`, i.e. channel = (z - z0)*(F*C) + f*C + c, channels-last."""
        B, H, W, ZC = rows_3d[0].shape
        planes = torch.stack([r.reshape(B, H, W, ZC // C, C) for r in rows_3d], dim=4)       # (B, H, W, z, frame, C)
        return planes[:, :, :, z0:z1].reshape(B, H, W, -1).contiguous()

    def _fuse_frames_acts(self, acts_2d, acts_3d):
        """fuse_frames on bf16 NHWC activations (channel = z*C + c per frame) -> the SFA's 512-channel input, every
        encoder writing its channel slice in place.  Channel orders as the reference's cat / unbind produce them: frames
        concatenated on C first, then z collapsed, i.e. channel = z*(F*C) + f*C + c."""
        from dhd_b200 import dense as D
        rows = lambda a: a.data.view(a.N, a.H, a.W, a.ld)[..., a.coff:a.coff + a.C]
        a0 = acts_3d[0]
        B, H, W, dev = a0.N, a0.H, a0.W, a0.data.device
        C = acts_2d[0].C                                                  # channels per z plane
        nz = a0.C // C
        x2 = self._frames_to_bev_rows([rows(a) for a in acts_2d])
        rows_3d = [rows(a) for a in acts_3d]
        enc = D.Act.empty(B, H, W, 512, 1, dev)
        feats = self.img_bev_encoder_backbone(D.Act(x2, x2.shape[-1], 1), return_act=True)
        self.img_bev_encoder_neck(feats, return_act=True, out=enc.slice(0, 256))
        lo = 256
        for i, (z0, z1) in enumerate(((0, 4), (4, 8), (8, nz))):
            slab = self._frames_to_slab_rows(rows_3d, C, z0, z1)
            net = getattr(self, 'img_voxel_encoder%d' % i)
            net(D.Act(slab, slab.shape[-1], 1), return_act=True, out=enc.slice(lo, lo + net.n_classes))
            lo += net.n_classes
        assert lo == 512
        return enc

    def _act_path_ok(self):
        """The inference fast path applies to the DHD-M / DHD-L wiring in the bf16 speed mode (every module of the
        chain hands bf16 NHWC activations to the next one)."""
        import os
        if os.environ.get('DHD_ACT_PATH', '1') == '0' or not getattr(self, 'act_path', True):
            return False
        if self.training or not self.pre_process or self.align_after_view_transfromation or not self.with_prev:
            return False
        vt = self.img_view_transformer
        mods = [vt, self.pre_process_net, self.pre_process_net_3d, self.img_bev_encoder_backbone, self.img_bev_encoder_neck,
                self.mix, self.occ_head] + [getattr(self, 'img_voxel_encoder%d' % i, None) for i in range(3)]
        if any(m is None or getattr(m, 'precision', None) != 'bf16' for m in mods) or getattr(vt, 'collapse_z', True):
            return False
        if any(type(getattr(self, 'img_voxel_neck%d' % i, None)).__name__ != 'Identity' for i in range(3)):
            return False
        widths = sum(getattr(self, 'img_voxel_encoder%d' % i).n_classes for i in range(3))
        return widths == 256 and getattr(self.mix, 'mix_channels', None) == 512

    def simple_test(self, points, img_metas, img=None, rescale=False, to_host=True, **kwargs):
        """DM:207-226; in the bf16 speed mode the chain pool -> pre-process nets -> encoders -> SFA -> head runs on bf16 NHWC
        activations end to end (`act_path = False` or DHD_ACT_PATH=0 selects the module-by-module tensor path)."""
        if not self._act_path_ok():
            return super().simple_test(points, img_metas, img=img, rescale=rescale, to_host=to_host, **kwargs)
        self._act_path = True
        try:
            with torch.no_grad():
                enc = self.extract_feat(points, img_inputs=img, img_metas=img_metas, **kwargs)[0]
                fused = self.mix(enc, return_act=True)
                occ = self.occ_head.forward_occ(fused)
                return self.occ_head.get_occ(occ, img_metas) if to_host else occ
        finally:
            self._act_path = False

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

    # ------------------------------------------------------------------ reference detector API (DM:377-666)
    def prepare_inputs(self, img_inputs, stereo=False):
        """bevdet4d.py:208-288: split the N = N_views * num_frame images / transforms into per-frame lists, derive the
        sensor -> key-ego transforms (fp64) and, for stereo, the current -> adjacent sensor transforms."""
        B, N = img_inputs[0].shape[:2]
        nf = self.num_frame
        N = N // nf
        imgs = img_inputs[0].view(B, N, nf, *img_inputs[0].shape[2:])
        imgs = [t.squeeze(2) for t in torch.split(imgs, 1, 2)]
        sensor2egos, ego2globals, intrins, post_rots, post_trans, bda = img_inputs[1:7]
        sensor2egos = sensor2egos.view(B, nf, N, 4, 4)
        ego2globals = ego2globals.view(B, nf, N, 4, 4)
        keyego2global = ego2globals[:, 0, 0, ...].unsqueeze(1).unsqueeze(1)
        global2keyego = torch.linalg.inv_ex(keyego2global.double()).inverse
        sensor2keyegos = (global2keyego @ ego2globals.double() @ sensor2egos.double()).float()
        curr2adjsensor = None
        if stereo:
            tf = self.temporal_frame
            s_curr, e_curr = sensor2egos[:, :tf].double(), ego2globals[:, :tf].double()
            s_adj, e_adj = sensor2egos[:, 1:tf + 1].double(), ego2globals[:, 1:tf + 1].double()
            c2a = (torch.linalg.inv_ex(e_adj @ s_adj).inverse @ e_curr @ s_curr).float()
            curr2adjsensor = [p.squeeze(1) for p in torch.split(c2a, 1, 1)] + [None] * self.extra_ref_frames
            assert len(curr2adjsensor) == nf
        extra = [sensor2keyegos, ego2globals, intrins.view(B, nf, N, 3, 3), post_rots.view(B, nf, N, 3, 3),
                 post_trans.view(B, nf, N, 3)]
        extra = [[p.squeeze(1) for p in torch.split(t, 1, 1)] for t in extra]
        sensor2keyegos, ego2globals, intrins, post_rots, post_trans = extra
        return imgs, sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda, curr2adjsensor

    def extract_stereo_ref_feat(self, x):
        """bevstereo4d.py:20-54: the first backbone stage of the extra reference frame (its 1/4 stereo feature only)."""
        if self.img_backbone is None or isinstance(self.img_backbone, MissingModule):
            raise NotImplementedError('extract_stereo_ref_feat needs the image backbone (outside the hot path): hand '
                                      'DHD_stereo the per-frame (features, stereo features) instead of images')
        return self.image_encoder(x, stereo=True)[1]

    def extract_img_feat(self, img_inputs, img_metas=None, pred_prev=False, sequential=False, **kwargs):
        """DM:377-541.  img_inputs[0]: (B, N_views * num_frame, 3, H, W) images -- or, with the image backbone outside
        this build, the pair (feats (B, N_views * num_frame, C, fH, fW), stereo_feats (B, N_views * num_frame, C_s, 4fH,
        4fW)) that image_encoder(img, stereo=True) would produce for every frame.
        -> (x_2d, x_3d, depth_key_frame, height_key_frame)."""
        if sequential or pred_prev:
            raise NotImplementedError('sequential / pred_prev inference (DM:401-402, 450-474) is not used by the DHD configs')
        first = img_inputs[0]
        injected = isinstance(first, (tuple, list))
        lead = first[0] if injected else first
        imgs, sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda, curr2adjsensor = \
            self.prepare_inputs([lead] + list(img_inputs[1:7]), stereo=True)
        if injected:
            B, NF = first[1].shape[:2]
            N = NF // self.num_frame
            st = first[1].view(B, N, self.num_frame, *first[1].shape[2:])
            stereo_feats = [t.squeeze(2).reshape(B * N, *first[1].shape[2:]) for t in torch.split(st, 1, 2)]
            feats = imgs
        else:
            feats, stereo_feats = [], []
            for fid, img in enumerate(imgs):
                if fid == self.num_frame - self.extra_ref_frames:
                    feats.append(None)
                    stereo_feats.append(self.extract_stereo_ref_feat(img))
                else:
                    x, sf = self.image_encoder(img, stereo=True)
                    feats.append(x)
                    stereo_feats.append(sf)
        return self.extract_bev_feat(feats, stereo_feats, sensor2keyegos, ego2globals, intrins, post_rots, post_trans,
                                     bda, curr2adjsensor)

    def extract_feat(self, points, img_inputs, img_metas=None, **kwargs):
        x_2d, x_3d, depth, height = self.extract_img_feat(img_inputs, img_metas, **kwargs)
        return x_2d, x_3d, None, depth, height

    def forward_train(self, points=None, img_metas=None, gt_bboxes_3d=None, gt_labels_3d=None, gt_labels=None,
                      gt_bboxes=None, img_inputs=None, proposals=None, gt_bboxes_ignore=None, **kwargs):
        """DM:577-614 -> dict(loss_depth, loss_height, loss_occ, loss_voxel_sem_scal, loss_voxel_geo_scal)."""
        x_2d, x_3d, _pts, depth, height = self.extract_feat(points, img_inputs=img_inputs, img_metas=img_metas, **kwargs)
        losses = dict()
        losses['loss_depth'], losses['loss_height'] = self.img_view_transformer.get_depth_and_height_loss(
            kwargs['gt_depth'], kwargs['gt_height'], depth, height)
        losses.update(self.forward_occ_train([x_2d, x_3d], kwargs['voxel_semantics'], kwargs['mask_camera']))
        return losses

    def forward_hot_path(self, feats, stereo_feats, sensor2keyegos, ego2globals, intrins, post_rots, post_trans, bda,
                         curr2adjsensor):
        """Per-frame image features -> occupancy logits (B, Dx, Dy, Dz, n_cls), key-frame depth and height."""
        x_2d, x_3d, depth, height = self.extract_bev_feat(feats, stereo_feats, sensor2keyegos, ego2globals, intrins,
                                                          post_rots, post_trans, bda, curr2adjsensor)
        fused = self.mix(torch.cat([x_2d, x_3d], dim=1), return_act=True)
        return self.occ_head(fused), depth, height
