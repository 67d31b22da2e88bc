"""Occupancy head `predictor` (reference: projects/mmdet3d_plugin/models/dense_heads/occ_head.py:32-153):
3x3 ConvModule (+ mmcv's default ReLU) -> permute to (B, Dx, Dy, C) -> Linear, Softplus, Linear ->
(B, Dx, Dy, Dz, n_cls).  Parameter names final_conv.conv.* / predicter.{0,2}.* as in the
reference; forward on tcgen05 GEMMs, the permute folded into the last layer's output strides."""
import numpy as np
import torch
import torch.nn as nn

from dhd_b200.compat import HEADS, BaseModule, ConvModule, EngineOwner, build_loss

# class frequencies of Occ3D-nuScenes used for the class-balanced CE weights (occ_head.py:10-29)
nusc_class_frequencies = np.array([
    944004, 1897170, 152386, 2391677, 16957802, 724139, 189027, 2074468, 413451, 2384460,
    5916653, 175883646, 4275424, 51393615, 61411620, 105975596, 116424404, 1892500630])


@HEADS.register_module(force=True)
class predictor(EngineOwner, BaseModule):
    def __init__(self, in_dim=256, out_dim=256, Dz=16, use_mask=True, weight_ce=1, weight_geo=1,
                 weight_sem=1, num_classes=18, use_predicter=True, class_balance=False, loss_occ=None,
                 precision='fp32'):
        super().__init__()
        self.in_dim, self.out_dim, self.Dz = in_dim, out_dim, Dz
        out_channels = out_dim if use_predicter else num_classes * Dz
        self.final_conv = ConvModule(in_dim, out_channels, kernel_size=3, stride=1, padding=1, bias=True,
                                     conv_cfg=dict(type='Conv2d'))
        self.use_predicter = use_predicter
        if use_predicter:
            self.predicter = nn.Sequential(nn.Linear(out_dim, out_dim * 2), nn.Softplus(),
                                           nn.Linear(out_dim * 2, num_classes * Dz))
        self.use_mask, self.num_classes, self.class_balance = use_mask, num_classes, class_balance
        if class_balance:
            self.cls_weights = torch.from_numpy(1 / np.log(nusc_class_frequencies[:num_classes] + 0.001))
            if loss_occ is not None:
                loss_occ = dict(loss_occ, class_weight=self.cls_weights)
        self.loss_occ = build_loss(loss_occ) if loss_occ is not None else None
        self.weight_ce, self.weight_geo, self.weight_sem = weight_ce, weight_geo, weight_sem
        self.precision = precision
        self._engine = None

    def _engine_for(self, img_feats):
        from dhd_b200 import dense as D
        from dhd_b200.modules import PredictorEngine
        if not isinstance(img_feats, D.Act):
            if not img_feats.is_cuda:
                raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
            img_feats = D.pack_any(img_feats, D.PRECISIONS[self.precision][0])
        dev = img_feats.data.device
        return self.cached_engine(dev, lambda: PredictorEngine(self, self.precision, dev)), img_feats

    def forward(self, img_feats):
        """img_feats: (B, C, Dy, Dx) fp32 CUDA tensor or a dhd_b200.dense.Act -> (B, Dx, Dy, Dz, n_cls).
        Under autograd the call runs the training engine and is differentiable (dhd_b200.autograd)."""
        from dhd_b200 import autograd as A
        from dhd_b200 import dense as D
        if not isinstance(img_feats, D.Act) and img_feats.is_cuda and A.wants_grad(self, img_feats):
            return A.predictor_forward(self, img_feats)
        with torch.no_grad():
            engine, img_feats = self._engine_for(img_feats)
            return engine(img_feats)

    def forward_occ(self, img_feats):
        """Inference tail in one go: forward + get_occ's softmax(-1).argmax(-1) (occ_head.py:141-153) -> uint8
        (B, Dx, Dy, Dz) on the device.  In the bf16 speed mode the class map comes out of the fused tail kernel's
        epilogue (dhd_predictor_tail) and the logits never reach HBM."""
        with torch.no_grad():
            engine, img_feats = self._engine_for(img_feats)
            occ = torch.empty(img_feats.N, img_feats.W, img_feats.H, self.Dz, dtype=torch.uint8,
                              device=img_feats.data.device)
            engine(img_feats, occ=occ, want_logits=False)
            return occ

    def loss(self, occ_pred, voxel_semantics, mask_camera):
        """occ_head.py:102-139: class-weighted masked cross-entropy + sem_scal + geo_scal (free class = 17), as a dict
        of differentiable torch scalars (works on any device).  The training engine of this package computes the same
        three terms and their gradient at the logits in one kernel pass (dhd_b200.train.PredictorTrainer.loss)."""
        from ..losses import geo_scal_loss_with_mask, sem_scal_loss_with_mask
        if not (self.use_mask and self.class_balance):
            raise NotImplementedError       # as the reference (occ_head.py:133-134)
        if self.loss_occ is None:
            raise RuntimeError('predictor.loss needs loss_occ (e.g. dict(type="CrossEntropyLoss", ...))')
        labels = voxel_semantics.long().reshape(-1)
        preds = occ_pred.reshape(-1, self.num_classes)
        mask = mask_camera.to(torch.int32).reshape(-1)
        cw = self.cls_weights.to(preds.device)
        valid = labels[mask.bool()]
        valid = valid[valid < self.num_classes]
        avg_factor = cw[valid].sum()                 # sum_i count(valid == i) * cls_weights[i], occ_head.py:114-117
        return dict(loss_occ=self.weight_ce * self.loss_occ(cls_score=preds, label=labels, weight=mask, avg_factor=avg_factor),
                    loss_voxel_sem_scal=self.weight_sem * sem_scal_loss_with_mask(preds, labels, mask),
                    loss_voxel_geo_scal=self.weight_geo * geo_scal_loss_with_mask(preds, labels, mask, non_empty_idx=17))

    def get_occ(self, occ_pred, img_metas=None):
        """(B, Dx, Dy, Dz, C) logits -> list of (Dx, Dy, Dz) uint8 class maps (occ_head.py:141-153).  CUDA logits go
        through dhd_occ_argmax (== softmax(-1).argmax(-1), tested incl. rounding ties); a uint8 tensor (the result of
        forward_occ) is passed through."""
        if occ_pred.dtype == torch.uint8:
            res = occ_pred
        elif occ_pred.is_cuda and occ_pred.dtype == torch.float32:
            from dhd_b200.modules import occ_argmax
            res = occ_argmax(occ_pred.detach().contiguous(),
                             torch.empty(occ_pred.shape[:-1], dtype=torch.uint8, device=occ_pred.device))
        else:
            res = occ_pred.softmax(-1).argmax(-1)
        return list(res.cpu().numpy().astype(np.uint8))
