"""Scene-class affinity terms of the occupancy loss (reference: models/losses/semkitti_loss.py:8-16, 136-225), masked
variants, as vectorised differentiable torch: the reference loops over the 17 classes with a `.sum() > 0` host sync in
every iteration; here per-class precision / recall / specificity come from three masked column sums.  Torch form for
autograd callers; the training path of this package evaluates the same terms inside dhd_occ_ce_loss."""
import torch
import torch.nn.functional as F

_EPS = 1e-5


def _bce_to_one(x):
    """binary_cross_entropy_with_logits(inverse_sigmoid(x), 1) = -log(x'), x' = x stepped off 0 / 1 by 1e-5 as the
    reference's inverse_sigmoid does (semkitti_loss.py:8-16)."""
    x = torch.where(x >= 1 - _EPS, x - _EPS, x)
    x = torch.where(x < _EPS, x + _EPS, x)
    return F.softplus(torch.log(1 / x - 1))


def _valid(ssc_target, mask, ignore_index):
    return ((ssc_target != ignore_index) & mask.bool()).to(torch.float32)


def geo_scal_loss_with_mask(pred, ssc_target, mask, ignore_index=255, non_empty_idx=0):
    """pred (N, C) logits, ssc_target (N,), mask (N,): occupied-vs-free precision / recall / specificity (136-168)."""
    p_free = F.softmax(pred, dim=1)[:, non_empty_idx]
    v = _valid(ssc_target, mask, ignore_index)
    occ_t = (ssc_target != non_empty_idx).to(torch.float32)
    occ_p = 1 - p_free
    hit = (occ_t * occ_p * v).sum()
    precision = hit / ((occ_p * v).sum() + _EPS)
    recall = hit / ((occ_t * v).sum() + _EPS)
    specificity = ((1 - occ_t) * p_free * v).sum() / (((1 - occ_t) * v).sum() + _EPS)
    return _bce_to_one(precision) + _bce_to_one(recall) + _bce_to_one(specificity)


def sem_scal_loss_with_mask(pred, ssc_target, mask, ignore_index=255):
    """Per-class precision / recall / specificity over classes 0 .. C-2, averaged over the classes present among the
    masked targets (170-225)."""
    n = pred.shape[1]
    v = _valid(ssc_target, mask, ignore_index)
    p = F.softmax(pred, dim=1) * v[:, None]
    tgt = F.one_hot(ssc_target.clamp(max=n - 1).long(), n).to(torch.float32) * ((ssc_target != ignore_index).float() * v)[:, None]
    cnt, sp, nom = tgt.sum(0)[:n - 1], p.sum(0)[:n - 1], (p * tgt).sum(0)[:n - 1]
    total = v.sum()
    rest = total - cnt                                          # masked voxels of the other classes
    zero = torch.zeros_like(cnt)
    terms = torch.where(sp > 0, _bce_to_one(nom / (sp + _EPS)), zero) + _bce_to_one(nom / (cnt + _EPS)) + \
        torch.where(rest > 0, _bce_to_one((rest - (sp - nom)) / (rest + _EPS)), zero)
    present = (cnt > 0).to(torch.float32)
    return (terms * present).sum() / present.sum()
