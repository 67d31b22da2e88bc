"""TEST INFRASTRUCTURE -- restatement of the occupancy head's loss terms (reference:
models/dense_heads/occ_head.py:102-131 and models/losses/semkitti_loss.py:8-16, 136-225) as vectorised,
differentiable torch (no Python loop over classes, no host syncs).  Pinned against the unmodified reference
functions (value and gradient) by tests/test_oracle.py where /root/reference exists."""
import torch
import torch.nn.functional as F

EPS = 1e-5


def _neg_log_clamped(x):
    """BCEWithLogits(inverse_sigmoid(x), 1) = -log(x') with semkitti_loss.inverse_sigmoid's stepping:
    x' = x - 1e-5 while x >= 1 - 1e-5, x + 1e-5 while x < 1e-5 (one step each for x in [0, 1])."""
    x = torch.where(x >= 1 - EPS, x - EPS, x)
    x = torch.where(x < EPS, x + EPS, x)
    return F.softplus(torch.log(1 / x - 1))


def ce_loss(preds, labels, mask, class_weight, ignore_index=255):
    """occ_head.py:112-127 with mmdet CrossEntropyLoss(class_weight), weight=mask, avg_factor=num_total_samples."""
    per = F.cross_entropy(preds, labels, weight=class_weight, reduction='none', ignore_index=ignore_index)
    valid = labels[mask.bool()]
    ok = valid != ignore_index
    avg = class_weight[valid[ok]].sum() if class_weight is not None else ok.sum().float()
    return (per * mask.float()).sum() / (avg + torch.finfo(torch.float32).eps)


def sem_scal_loss_with_mask(preds, labels, mask, ignore_index=255):
    """semkitti_loss.py:170-225: classes 0..n-2, averaged over the classes present in the masked target."""
    p = F.softmax(preds, dim=1)
    m = ((labels != ignore_index) & mask.bool()).float()
    n = p.shape[1]
    onehot = F.one_hot(labels.clamp(max=n - 1).long(), n).float() * (labels != ignore_index).float()[:, None]
    c = onehot * m[:, None]                       # completion target of every class on the masked voxels
    pm = p * m[:, None]
    cnt = c.sum(0)
    nom = (pm * c).sum(0)
    sp = pm.sum(0)
    M = m.sum()
    loss = preds.new_zeros(())
    present = cnt[:n - 1] > 0
    for_all = lambda t: t[:n - 1]
    prec = for_all(nom) / (for_all(sp) + EPS)
    rec = for_all(nom) / (for_all(cnt) + EPS)
    spec_den = M - for_all(cnt)
    spec = ((M - for_all(cnt)) - (for_all(sp) - for_all(nom))) / (spec_den + EPS)
    terms = torch.where(for_all(sp) > 0, _neg_log_clamped(prec), torch.zeros_like(prec)) + _neg_log_clamped(rec) + \
        torch.where(spec_den > 0, _neg_log_clamped(spec), torch.zeros_like(spec))
    loss = (terms * present.float()).sum()
    return loss / present.float().sum()


def geo_scal_loss_with_mask(preds, labels, mask, ignore_index=255, non_empty_idx=17):
    """semkitti_loss.py:136-168."""
    p = F.softmax(preds, dim=1)
    m = ((labels != ignore_index) & mask.bool()).float()
    empty = p[:, non_empty_idx]
    nonempty = 1 - empty
    t = (labels != non_empty_idx).float()
    inter = (t * nonempty * m).sum()
    prec = inter / ((nonempty * m).sum() + EPS)
    rec = inter / ((t * m).sum() + EPS)
    spec = ((1 - t) * empty * m).sum() / (((1 - t) * m).sum() + EPS)
    return _neg_log_clamped(prec) + _neg_log_clamped(rec) + _neg_log_clamped(spec)


def predictor_loss(preds, labels, mask, class_weight, weight_ce=1.0, weight_sem=1.0, weight_geo=1.0):
    """The three terms of predictor.loss (occ_head.py:124-131) as a dict."""
    return dict(loss_occ=weight_ce * ce_loss(preds, labels, mask, class_weight),
                loss_voxel_sem_scal=weight_sem * sem_scal_loss_with_mask(preds, labels, mask),
                loss_voxel_geo_scal=weight_geo * geo_scal_loss_with_mask(preds, labels, mask, non_empty_idx=17))


def _min_pool_bins(gt, ds, lo, interval, nbins):
    """LH:625-701 (get_downsampled_gt_depth / _height, sid=False): ds x ds min-pool of a sparse (B, N, H, W) map with
    zeros ignored (1e5 sentinel), binned; one-hot over nbins + 1 classes with class 0 (= out of range) dropped."""
    B, N, H, W = gt.shape
    t = torch.where(gt == 0.0, torch.full_like(gt, 1e5), gt)
    t = t.view(B * N, H // ds, ds, W // ds, ds).permute(0, 1, 3, 2, 4).reshape(-1, ds * ds).min(dim=-1).values
    t = (t - lo) / interval
    t = torch.where((t < nbins + 1) & (t >= 0.0), t, torch.zeros_like(t))
    return F.one_hot(t.long(), num_classes=nbins + 1).view(-1, nbins + 1)[:, 1:].float()


def height_loss(gt_depth, gt_height, height, depth_cfg, n_depth, height_lo, height_interval, weight, downsample=16):
    """MGHS.get_height_loss (LH:595-622): BCE between the height distribution (B*N, H, fH, fW) and the binned LiDAR
    height on the pixels whose LiDAR depth falls into a depth bin, / max(1, #foreground), x loss_height_weight.
    depth_cfg: the [lo, hi, step] the module holds at loss time (mask_3_grid['depth'] for MGHS, the LH:455 quirk)."""
    H = height.shape[1]
    labels = _min_pool_bins(gt_height, downsample, height_lo, height_interval, H)
    fg = _min_pool_bins(gt_depth, downsample, depth_cfg[0] - depth_cfg[2], depth_cfg[2], n_depth).max(dim=1).values > 0.0
    preds = height.permute(0, 2, 3, 1).contiguous().view(-1, H)
    loss = F.binary_cross_entropy(preds[fg].float(), labels[fg], reduction='none').sum()
    return weight * loss / max(1.0, float(fg.sum()))
