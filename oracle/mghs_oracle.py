"""TEST INFRASTRUCTURE -- CPU oracle for the MGHS view transform and bev_pool_v2.

A restatement (not a copy) of the reference algorithm, written against the
reference's own torch calls wherever summation order matters so that on CPU it is
bit-identical to the reference Python.  Pinned against the real reference code by
``tests/test_oracle_vs_reference.py`` (runs where /root/reference exists) and
against the committed fixtures in ``tests/golden/`` everywhere else.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product path
(``dhd_b200/``, ``projects/``) never does.

LH = projects/mmdet3d_plugin/models/necks/lss_heightmap.py
BP = projects/mmdet3d_plugin/ops/bev_pool_v2/
"""
import ctypes
import math
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- grids
def grid_infos(x, y, z):
    """LH:86-102 -- lower bound / interval / size as fp32 tensors built from Python
    floats: size = (hi - lo) / step evaluated in float64 *then* rounded to fp32."""
    lower = torch.tensor([c[0] for c in (x, y, z)], dtype=torch.float32)
    interval = torch.tensor([c[2] for c in (x, y, z)], dtype=torch.float32)
    size = torch.tensor([(c[1] - c[0]) / c[2] for c in (x, y, z)], dtype=torch.float32)
    return lower, interval, size


def frustum(depth_cfg, input_size, downsample):
    """LH:105-134 (sid=False): (D, fH, fW, 3) template of (u, v, d)."""
    h_in, w_in = input_size
    fh, fw = h_in // downsample, w_in // downsample
    d = torch.arange(*depth_cfg, dtype=torch.float32)
    u = torch.linspace(0, w_in - 1, fw, dtype=torch.float32)
    v = torch.linspace(0, h_in - 1, fh, dtype=torch.float32)
    D = d.shape[0]
    out = torch.empty(D, fh, fw, 3, dtype=torch.float32)
    out[..., 0] = u.view(1, 1, fw)
    out[..., 1] = v.view(1, fh, 1)
    out[..., 2] = d.view(D, 1, 1)
    return out


# ----------------------------------------------------------------------- geometry
def ego_coor(frus, sensor2ego, cam2imgs, post_rots, post_trans, bda):
    """LH:179-231 -- frustum points in the ego frame, (B, N, D, fH, fW, 3).

    Same torch calls in the same order as the reference so the broadcast batched
    matmuls take the same backend path."""
    B, N = sensor2ego.shape[:2]
    pts = frus.to(sensor2ego) - post_trans.view(B, N, 1, 1, 1, 3)
    pts = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1))
    pts = torch.cat((pts[..., :2, :] * pts[..., 2:3, :], pts[..., 2:3, :]), 5)
    comb = sensor2ego[:, :, :3, :3].matmul(torch.inverse(cam2imgs))
    pts = comb.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1)
    pts = pts + sensor2ego[:, :, :3, 3].view(B, N, 1, 1, 1, 3)
    pts = bda.view(B, 1, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1)).squeeze(-1)
    return pts


def camera_matrices(sensor2ego, cam2imgs, post_rots, post_trans, bda):
    """The per-camera 3x3s the reference derives with torch (LH:206, 217): these are
    the *inputs* of the fused CUDA geometry kernel.  Returns fp32 contiguous
    (inv_post_rot (BN,3,3), post_tran (BN,3), combine (BN,3,3), trans (BN,3), bda (B,3,3))."""
    B, N = sensor2ego.shape[:2]
    ipr = torch.inverse(post_rots).reshape(B * N, 3, 3).contiguous()
    comb = sensor2ego[:, :, :3, :3].matmul(torch.inverse(cam2imgs)).reshape(B * N, 3, 3).contiguous()
    tr = sensor2ego[:, :, :3, 3].reshape(B * N, 3).contiguous()
    return ipr, post_trans.reshape(B * N, 3).contiguous(), comb, tr, bda.contiguous()


def quantise(coor, lower, interval, size):
    """LH:331-342 -- subtract-then-divide in fp32, truncate toward zero, keep test.
    Returns (idx int64 (...,3), kept bool (...))."""
    g = (coor - lower.to(coor)) / interval.to(coor)
    idx = g.long()
    kept = ((idx[..., 0] >= 0) & (idx[..., 0] < size[0]) &
            (idx[..., 1] >= 0) & (idx[..., 1] < size[1]) &
            (idx[..., 2] >= 0) & (idx[..., 2] < size[2]))
    return idx, kept


def prepare_v2(coor, lower, interval, size):
    """LH:303-371 -- ranks + run-length intervals.  Sort is made STABLE here (the
    reference's argsort is unstable, so only the per-interval *sets* are defined)."""
    B, N, D, H, W, _ = coor.shape
    npts = B * N * D * H * W
    ranks_depth = torch.arange(npts, dtype=torch.int32)
    ranks_feat = torch.arange(npts // D, dtype=torch.int32).reshape(B, N, 1, H, W)
    ranks_feat = ranks_feat.expand(B, N, D, H, W).flatten()
    idx, kept = quantise(coor, lower, interval, size)
    idx = idx.view(npts, 3)
    kept = kept.view(npts)
    batch = torch.arange(B).view(B, 1).expand(B, npts // B).reshape(npts)
    if npts == 0:
        return None, None, None, None, None
    idx, batch = idx[kept], batch[kept]
    ranks_depth, ranks_feat = ranks_depth[kept], ranks_feat[kept]
    # fp32 rank arithmetic, as the reference (grid_size is a float tensor), LH:351-354
    rb = batch.to(torch.float32) * (size[2] * size[1] * size[0])
    rb = rb + idx[:, 2].to(torch.float32) * (size[1] * size[0])
    rb = rb + (idx[:, 1].to(torch.float32) * size[0] + idx[:, 0].to(torch.float32))
    order = torch.argsort(rb, stable=True)
    rb, ranks_depth, ranks_feat = rb[order], ranks_depth[order], ranks_feat[order]
    first = torch.ones(rb.shape[0], dtype=torch.bool)
    first[1:] = rb[1:] != rb[:-1]
    starts = torch.where(first)[0].int()
    if len(starts) == 0:
        return None, None, None, None, None
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = rb.shape[0] - starts[-1]
    return (rb.int().contiguous(), ranks_depth.int().contiguous(),
            ranks_feat.int().contiguous(), starts.contiguous(), lengths.contiguous())


# --------------------------------------------------------------- bev_pool_v2 (CPU)
_clib = None


def _c_oracle():
    """The plain-C restatement of the two CUDA kernels (oracle/bev_pool_ref.c)."""
    global _clib
    if _clib is None:
        so = os.path.join(_HERE, '_build', 'libdhd_oracle.so')
        if not os.path.exists(so):
            from .build_oracle import build
            build()
        _clib = ctypes.CDLL(so)
        _clib.oracle_bev_pool_v2_fwd.restype = None
        _clib.oracle_bev_pool_v2_bwd.restype = None
        _clib.oracle_set_threads.restype = ctypes.c_int
    return _clib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def pool_fwd_c(depth, feat, rd, rf, rb, starts, lengths, out, threads=1):
    """out[rb[start]] = sum_i depth[rd_i] * feat[rf_i, :]  (BP/src/bev_pool_cuda.cu:21-50)."""
    lib = _c_oracle()
    lib.oracle_set_threads(int(threads))
    c = feat.shape[-1]
    lib.oracle_bev_pool_v2_fwd(ctypes.c_int(c), ctypes.c_int(starts.numel()), _p(depth), _p(feat),
                               _p(rd), _p(rf), _p(rb), _p(starts), _p(lengths), _p(out))
    return out


def pool_bwd_c(out_grad, depth, feat, rd, rf, rb, starts, lengths, depth_grad, feat_grad, threads=1):
    """BP/src/bev_pool_cuda.cu:69-123 with intervals over the ranks_feat-sorted points."""
    lib = _c_oracle()
    lib.oracle_set_threads(int(threads))
    c = feat.shape[-1]
    lib.oracle_bev_pool_v2_bwd(ctypes.c_int(c), ctypes.c_int(starts.numel()), _p(out_grad), _p(depth),
                               _p(feat), _p(rd), _p(rf), _p(rb), _p(starts), _p(lengths),
                               _p(depth_grad), _p(feat_grad))
    return depth_grad, feat_grad


def _runs(keys):
    first = torch.ones(keys.shape[0], dtype=torch.bool)
    first[1:] = keys[1:] != keys[:-1]
    starts = torch.where(first)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = keys.shape[0] - starts[-1]
    return starts, lengths


class _PoolFn(torch.autograd.Function):
    """BP/bev_pool.py:11-83 (QuickCumsumCuda) on CPU through the C restatement."""

    @staticmethod
    def forward(ctx, depth, feat, rd, rf, rb, shape, starts, lengths):
        depth = depth.contiguous().float()
        feat = feat.contiguous().float()
        rd, rf, rb = rd.contiguous().int(), rf.contiguous().int(), rb.contiguous().int()
        starts, lengths = starts.contiguous().int(), lengths.contiguous().int()
        out = feat.new_zeros(shape)
        pool_fwd_c(depth, feat, rd, rf, rb, starts, lengths, out, threads=_PoolFn.threads)
        ctx.save_for_backward(rb, depth, feat, rf, rd)
        return out

    @staticmethod
    def backward(ctx, g):
        rb, depth, feat, rf, rd = ctx.saved_tensors
        order = torch.argsort(rf, stable=True)          # BP/bev_pool.py:47-49
        rf, rd, rb = rf[order].contiguous(), rd[order].contiguous(), rb[order].contiguous()
        starts, lengths = _runs(rf)                      # BP/bev_pool.py:50-57
        dg = depth.new_zeros(depth.shape)
        fg = feat.new_zeros(feat.shape)
        pool_bwd_c(g.contiguous(), depth, feat, rd, rf, rb, starts, lengths, dg, fg,
                   threads=_PoolFn.threads)
        return dg, fg, None, None, None, None, None, None

    threads = 1


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                interval_starts, interval_lengths):
    """BP/bev_pool.py:86-106 -- returns (B, C, Dz, Dy, Dx) contiguous."""
    x = _PoolFn.apply(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                      interval_starts, interval_lengths)
    return x.permute(0, 4, 1, 2, 3).contiguous()


def bev_pool_v2_index_add(depth, feat, rd, rf, rb, shape):
    """Independent second restatement (pure torch index_add_) used to cross-check the C one."""
    c = feat.shape[-1]
    out = torch.zeros(shape, dtype=torch.float32).view(-1, c)
    out.index_add_(0, rb.long(), depth.flatten()[rd.long(), None] * feat.reshape(-1, c)[rf.long()])
    return out.view(shape).permute(0, 4, 1, 2, 3).contiguous()


# -------------------------------------------------------------------- height masks
def height_masks(height, height_range, mask_range):
    """LH:528-564 -- argmax bin -> fp32 table value -> three half-open masks.
    Returns (mask_id int8 (BN,fH,fW): 0 = none, 1/2/3 = low/mid/high, and the 3 bool masks)."""
    k = torch.argmax(height, dim=1)
    hv = torch.tensor(height_range, device=height.device)[k]
    h_min, t1, t2, h_max = mask_range
    m1 = (hv >= h_min) & (hv < t1)
    m2 = (hv >= t1) & (hv < t2)
    m3 = (hv >= t2) & (hv < h_max)
    mid = m1.to(torch.int8) + 2 * m2.to(torch.int8) + 3 * m3.to(torch.int8)
    return mid, (m1, m2, m3)


# ----------------------------------------------------------------- view transform
BEV_GRID = {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [-1, 5.4, 6.4]}   # LH:425-431


def pool_one_pass(coor, depth, feat_nchw, grid, collapse_z=True):
    """LH:261-300 for one grid.  depth (B,N,D,fH,fW); feat (B,N,C,fH,fW)."""
    lower, interval, size = grid_infos(grid['x'], grid['y'], grid['z'])
    rb, rd, rf, st, ln = prepare_v2(coor, lower, interval, size)
    B = depth.shape[0]
    C = feat_nchw.shape[2]
    shape = (B, int(size[2]), int(size[1]), int(size[0]), C)
    feat = feat_nchw.permute(0, 1, 3, 4, 2)
    if rb is None:
        out = torch.zeros(shape[0], C, shape[1], shape[2], shape[3])
    else:
        out = bev_pool_v2(depth, feat, rd, rf, rb, shape, st, ln)
    if collapse_z:
        out = torch.cat(out.unbind(dim=2), 1)
    return out


def view_transform(inputs, depth, tran_feat, height, frus, height_range, mask_range,
                   mask_grids, collapse_z=True, bev_grid=BEV_GRID):
    """LH:407-459 (MGHS, collapse_z=True) / LH:793-856 (MGHS_Depth: pass collapse_z=False and
    concatenate L/M/H on z yourself).  inputs = (x, sensor2ego, ego2global, K, post_rot,
    post_tran, bda).  Returns (bev, L, M, H)."""
    x, s2e, _e2g, K, pr, pt, bda = inputs[:7]
    B, N, _, fH, fW = x.shape
    D = depth.shape[1]
    C = tran_feat.shape[1]
    coor = ego_coor(frus, s2e, K, pr, pt, bda)
    d5 = depth.view(B, N, D, fH, fW)
    outs = [pool_one_pass(coor, d5, tran_feat.view(B, N, C, fH, fW), bev_grid, collapse_z)]
    _, masks = height_masks(height, height_range, mask_range)
    for m, g in zip(masks, mask_grids):
        mf = tran_feat * m.unsqueeze(1).expand_as(tran_feat)
        outs.append(pool_one_pass(coor, d5, mf.view(B, N, C, fH, fW), g, collapse_z))
    return tuple(outs)


# ------------------------------------------------------------------- synthetic rig
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

CFG1 = dict(   # BASELINE.json configs[0]: the CPU plumbing case, one pass, no masks
    input_size=(64, 176), downsample=16, depth=[1.0, 60.0, 1.0], C=64, C_in=64,
    height_range=None, mask_range=None, mask_grids=[],
    bev_grid={'x': [-40, 40, 1.6], 'y': [-40, 40, 1.6], 'z': [-1, 5.4, 1.6]}, ncams=1,
)

MINI = dict(   # DHD-S passes / masks on a coarse 50x50 grid and a small image: full outputs fit
    input_size=(64, 176), downsample=16, depth=[1.0, 45.0, 1.0], C=64, C_in=64,
    height_range=DHD_S['height_range'], mask_range=DHD_S['mask_range'],
    mask_grids=[dict(g, x=[-40, 40, 1.6], y=[-40, 40, 1.6]) for g in DHD_S['mask_grids']],
    bev_grid=BEV_GRID,   # MGHS.view_transform hard-codes the BEV pass grid (LH:425-431)
    ncams=2,
)


def synthetic_rig(B, ncams=6, input_size=(256, 704), src_size=(900, 1600), seed=0, flip_bda=False):
    """6 surround cameras, nuScenes-like intrinsics, resize+crop aug."""
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
    pt[..., 1] = -(src_size[0] * scale - input_size[0])       # crop the top rows
    bda = torch.eye(3).expand(B, 3, 3).clone()
    if flip_bda:
        bda[1::2, 0, 0] = -1.0
        bda[1::2, 1, 1] = -1.0
    e2g = torch.eye(4).expand(B, ncams, 4, 4).clone()
    return s2e, e2g, K, pr, pt, bda


def synthetic_inputs(cfg, B, seed=0, flip_bda=False, rig=None):
    """Seeded synthetic tensors for one view-transform call.

    Values are integers scaled by powers of two (exactly representable, no libm / vectorised
    transcendental in the generator) so they are bit-identical on every host: depth in
    (0, 1], context in [-4, 4), height scores in [0, 1).  depth / height are "already
    softmaxed" as far as the view transform is concerned (it never re-normalises them)."""
    g = torch.Generator().manual_seed(seed + 1)
    N = cfg['ncams']
    h_in, w_in = cfg['input_size']
    fH, fW = h_in // cfg['downsample'], w_in // cfg['downsample']
    D = torch.arange(*cfg['depth']).shape[0]
    if rig is None:
        rig = synthetic_rig(B, N, cfg['input_size'], seed=seed, flip_bda=flip_bda)
    x = torch.zeros(B, N, 1, fH, fW)        # only its shape is read by view_transform
    depth = torch.randint(1, 1025, (B * N, D, fH, fW), generator=g).float() * (1.0 / 1024.0)
    feat = torch.randint(-2048, 2048, (B * N, cfg['C'], fH, fW), generator=g).float() * (1.0 / 512.0)
    height = None
    if cfg['height_range'] is not None:
        height = torch.randint(0, 1 << 20, (B * N, len(cfg['height_range']), fH, fW),
                               generator=g).float() * (1.0 / (1 << 20))
    return (x,) + tuple(rig), depth, feat, height
