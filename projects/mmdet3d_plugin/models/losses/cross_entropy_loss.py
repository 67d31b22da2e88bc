"""`CrossEntropyLoss` of the LOSSES registry (reference: models/losses/cross_entropy_loss.py, the mmdet loss the
occupancy head is configured with: `loss_occ=dict(type='CrossEntropyLoss', use_sigmoid=False, ignore_index=255,
loss_weight=1.0)`, DHD-S.py occ_head).  Softmax mode only -- the sigmoid / mask modes are not used by any DHD config.

This is the torch form for callers that want autograd; the training path of this package computes the same value and
its gradient in one kernel (dhd_occ_ce_loss, dhd_b200.train.PredictorTrainer.loss)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from dhd_b200.compat import LOSSES


@LOSSES.register_module(force=True)
class CrossEntropyLoss(nn.Module):
    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None, ignore_index=None,
                 loss_weight=1.0, avg_non_ignore=False):
        super().__init__()
        if use_sigmoid or use_mask:
            raise NotImplementedError('CrossEntropyLoss: only the softmax mode of the DHD configs is built')
        self.use_sigmoid, self.use_mask = use_sigmoid, use_mask
        self.reduction, self.loss_weight, self.class_weight = reduction, loss_weight, class_weight
        self.ignore_index, self.avg_non_ignore = ignore_index, avg_non_ignore

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=None,
                **kwargs):
        """cls_score (N, C) logits, label (N,), weight (N,) per-sample weight (the camera mask), avg_factor: divisor of
        the weighted sum (occ_head.py:112-127 passes the class-weighted count of masked voxels)."""
        reduction = reduction_override or self.reduction
        ignore_index = self.ignore_index if ignore_index is None else ignore_index
        ignore_index = -100 if ignore_index is None else ignore_index
        cw = None
        if self.class_weight is not None:
            cw = torch.as_tensor(self.class_weight, dtype=cls_score.dtype, device=cls_score.device)
        loss = F.cross_entropy(cls_score, label, weight=cw, reduction='none', ignore_index=ignore_index)
        if avg_factor is None and self.avg_non_ignore and reduction == 'mean':
            avg_factor = label.numel() - (label == ignore_index).sum().item()
        if weight is not None:
            loss = loss * weight.float()
        if avg_factor is None:
            loss = loss.mean() if reduction == 'mean' else (loss.sum() if reduction == 'sum' else loss)
        elif reduction == 'mean':
            loss = loss.sum() / (avg_factor + torch.finfo(torch.float32).eps)     # mmdet weight_reduce_loss
        elif reduction != 'none':
            raise ValueError('avg_factor can not be used with reduction="sum"')
        return self.loss_weight * loss
