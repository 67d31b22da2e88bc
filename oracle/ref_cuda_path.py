"""TEST / MEASUREMENT INFRASTRUCTURE -- the reference's CUDA path of MGHS.view_transform.

The reference runs, per forward and per grid (4 grids for DHD-S), on the GPU:
  get_ego_coor (LH:179-231)  ->  voxel_pooling_prepare_v2 (LH:303-371: quantise, filter,
  fp32 ranks, argsort, run-length)  ->  new_zeros + bev_pool_v2 kernel (BP/bev_pool.py:17-41,
  BP/src/bev_pool_cuda.cu:21-50)  ->  permute(0,4,1,2,3).contiguous() (BP/bev_pool.py:105)
  ->  cat(unbind(2), 1) (LH:298-299),  plus the three masked feature copies (LH:436-442).
This module restates exactly that op sequence with torch CUDA ops and launches the reference's
OWN kernel, compiled unmodified from /root/reference into oracle/_ref/libbev_pool_v2_ref.so
(oracle/build_oracle.py).  It is the "reference bev_pool_v2 CUDA path" of BASELINE.json's
north_star: the denominator of the >=10x target and a full-size second checker.

LH = projects/mmdet3d_plugin/models/necks/lss_heightmap.py, BP = .../ops/bev_pool_v2/.
Only tests/ and scripts/bench_ref_cuda.py import this; the product path never does.
"""
import ctypes
import os

import torch

from . import mghs_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, '_ref', 'libbev_pool_v2_ref.so')
_ref = None
_ref_grad = None


def available():
    return os.path.exists(REF_SO)


def _kernel():
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(REF_SO)
        fn = getattr(lib, '_Z11bev_pool_v2iiPKfS0_PKiS2_S2_S2_S2_Pf')   # void bev_pool_v2(int c, int n_intervals, ...)
        fn.restype = None
        _ref = fn
    return _ref


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def bev_pool_v2_ref(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                    interval_lengths):
    """BP/bev_pool.py:17-41 + 86-106 with the reference kernel (legacy default stream, as the
    reference launches it: bev_pool_cuda.cu:129)."""
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    out = feat.new_zeros(bev_feat_shape)
    if torch.cuda.current_stream() != torch.cuda.default_stream():
        torch.cuda.current_stream().synchronize()     # the reference kernel runs on the legacy default stream
    _kernel()(ctypes.c_int(feat.shape[-1]), ctypes.c_int(interval_starts.numel()), _p(depth), _p(feat),
              _p(ranks_depth), _p(ranks_feat), _p(ranks_bev), _p(interval_starts), _p(interval_lengths), _p(out))
    return out.permute(0, 4, 1, 2, 3).contiguous()


def _grad_kernel():
    global _ref_grad
    if _ref_grad is None:
        lib = ctypes.CDLL(REF_SO)
        fn = getattr(lib, '_Z16bev_pool_v2_gradiiPKfS0_S0_PKiS2_S2_S2_S2_PfS3_')   # void bev_pool_v2_grad(int c, int n_intervals, ...)
        fn.restype = None
        _ref_grad = fn
    return _ref_grad


def bev_pool_v2_grad_ref(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev):
    """QuickCumsumCuda.backward (BP/bev_pool.py:44-83) with the reference's own bev_pool_v2_grad launcher
    (BP/src/bev_pool_cuda.cu:69-123, 135-142): re-sort the points by ranks_feat, run-length the feature
    intervals, zero-filled depth_grad / feat_grad, one thread per feature interval.
    out_grad: (B, Dz, Dy, Dx, C) -- the gradient of the kernel's own output layout."""
    order = ranks_feat.argsort()
    ranks_feat, ranks_depth, ranks_bev = ranks_feat[order], ranks_depth[order], ranks_bev[order]
    kept = torch.ones(ranks_bev.shape[0], device=ranks_bev.device, dtype=torch.bool)
    kept[1:] = ranks_feat[1:] != ranks_feat[:-1]
    starts = torch.where(kept)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = ranks_bev.shape[0] - starts[-1]
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    depth_grad = depth.new_zeros(depth.shape)
    feat_grad = feat.new_zeros(feat.shape)
    out_grad = out_grad.contiguous().float()
    rd, rf, rb = ranks_depth.contiguous(), ranks_feat.contiguous(), ranks_bev.contiguous()
    starts, lengths = starts.contiguous(), lengths.contiguous()
    torch.cuda.synchronize()                          # the reference kernel runs on the legacy default stream
    _grad_kernel()(ctypes.c_int(feat.shape[-1]), ctypes.c_int(starts.numel()), _p(out_grad), _p(depth), _p(feat),
                   _p(rd), _p(rf), _p(rb), _p(starts), _p(lengths), _p(depth_grad), _p(feat_grad))
    torch.cuda.synchronize()
    return depth_grad, feat_grad


def prepare_v2_cuda(coor, lower, interval, size):
    """LH:303-371 on the device of `coor` (unstable argsort, fp32 rank arithmetic)."""
    B, N, D, H, W, _ = coor.shape
    num_points = B * N * D * H * W
    dev = coor.device
    ranks_depth = torch.arange(0, num_points, dtype=torch.int, device=dev)
    ranks_feat = torch.arange(0, num_points // D, dtype=torch.int, device=dev)
    ranks_feat = ranks_feat.reshape(B, N, 1, H, W).expand(B, N, D, H, W).flatten()
    coor = ((coor - lower.to(coor)) / interval.to(coor))
    coor = coor.long().view(num_points, 3)
    batch_idx = torch.arange(0, B, dtype=torch.float32).reshape(B, 1).expand(B, num_points // B) \
        .reshape(num_points, 1).to(coor)
    coor = torch.cat((coor, batch_idx), 1)
    kept = (coor[:, 0] >= 0) & (coor[:, 0] < size[0]) & (coor[:, 1] >= 0) & (coor[:, 1] < size[1]) & \
           (coor[:, 2] >= 0) & (coor[:, 2] < size[2])
    coor, ranks_depth, ranks_feat = coor[kept], ranks_depth[kept], ranks_feat[kept]
    ranks_bev = coor[:, 3] * (size[2] * size[1] * size[0])
    ranks_bev += coor[:, 2] * (size[1] * size[0])
    ranks_bev += coor[:, 1] * size[0] + coor[:, 0]
    order = ranks_bev.argsort()
    ranks_bev, ranks_depth, ranks_feat = ranks_bev[order], ranks_depth[order], ranks_feat[order]
    kept = torch.ones(ranks_bev.shape[0], device=dev, dtype=torch.bool)
    kept[1:] = ranks_bev[1:] != ranks_bev[:-1]
    interval_starts = torch.where(kept)[0].int()
    interval_lengths = torch.zeros_like(interval_starts)
    interval_lengths[:-1] = interval_starts[1:] - interval_starts[:-1]
    interval_lengths[-1] = ranks_bev.shape[0] - interval_starts[-1]
    return (ranks_bev.int().contiguous(), ranks_depth.int().contiguous(), ranks_feat.int().contiguous(),
            interval_starts.int().contiguous(), interval_lengths.int().contiguous())


def view_transform_core_cuda(inputs, depth5, feat_nchw, frus, grid, collapse_z=True):
    """LH:380-405 (view_transform_core) + 261-300 for one grid, everything on the GPU."""
    _x, s2e, _e2g, K, pr, pt, bda = inputs[:7]
    coor = O.ego_coor(frus, s2e, K, pr, pt, bda)                 # recomputed per pass, as the reference does
    lower, interval, size = O.grid_infos(grid['x'], grid['y'], grid['z'])
    size = size.to(coor.device)                                   # the reference keeps grid_size where it was made
    rb, rd, rf, st, ln = prepare_v2_cuda(coor, lower, interval, size)
    B, C = depth5.shape[0], feat_nchw.shape[2]
    feat = feat_nchw.permute(0, 1, 3, 4, 2)
    shape = (B, int(size[2]), int(size[1]), int(size[0]), C)
    out = bev_pool_v2_ref(depth5, feat, rd, rf, rb, shape, st, ln)
    if collapse_z:
        out = torch.cat(out.unbind(dim=2), 1)
    return out


def view_transform_cuda(inputs, depth, tran_feat, height, frus, height_range, mask_range, mask_grids,
                        collapse_z=True, bev_grid=O.BEV_GRID):
    """LH:407-459: BEV pass + the three height-masked passes; returns (bev, L, M, H)."""
    x = inputs[0]
    B, N, _, fH, fW = x.shape
    D, C = depth.shape[1], tran_feat.shape[1]
    d5 = depth.view(B, N, D, fH, fW)
    outs = [view_transform_core_cuda(inputs, d5, tran_feat.view(B, N, C, fH, fW), frus, bev_grid, collapse_z)]
    _, masks = O.height_masks(height, height_range, mask_range)
    for m, g in zip(masks, mask_grids):
        mf = tran_feat * m.unsqueeze(1).expand_as(tran_feat)
        outs.append(view_transform_core_cuda(inputs, d5, mf.view(B, N, C, fH, fW), frus, g, collapse_z))
    return tuple(outs)


def view_transform_backward_cuda(inputs, depth, tran_feat, height, frus, height_range, mask_range, mask_grids,
                                 out_grads, collapse_z=True, bev_grid=O.BEV_GRID):
    """What autograd does for the reference's four passes (LH:407-459 -> BP/bev_pool.py:44-83): for every grid the
    reference's grad kernel on that pass's ranks; depth receives the SUM of the four depth gradients (one softmax
    output feeds all passes), tran_feat the BEV-pass gradient plus mask_k x the k-th masked pass's gradient
    (d(tran_feat * mask)/d tran_feat = mask, LH:436-442).
    out_grads: gradients of view_transform_cuda()'s outputs -- (B, Dz*C, Dy, Dx) when collapse_z else
    (B, C, Dz, Dy, Dx).  Returns (depth_grad (B*N, D, fH, fW), feat_grad (B*N, C, fH, fW))."""
    x = inputs[0]
    _x, s2e, _e2g, K, pr, pt, bda = inputs[:7]
    B, N, _, fH, fW = x.shape
    D, C = depth.shape[1], tran_feat.shape[1]
    d5 = depth.view(B, N, D, fH, fW)
    _, masks = O.height_masks(height, height_range, mask_range)
    dgrad = torch.zeros_like(d5)
    fgrad = torch.zeros(B * N, C, fH, fW, device=depth.device)
    for p, grid in enumerate([bev_grid] + list(mask_grids)):
        coor = O.ego_coor(frus, s2e, K, pr, pt, bda)
        lower, interval, size = O.grid_infos(grid['x'], grid['y'], grid['z'])
        size = size.to(coor.device)
        rb, rd, rf, _st, _ln = prepare_v2_cuda(coor, lower, interval, size)
        feat = tran_feat if p == 0 else tran_feat * masks[p - 1].unsqueeze(1).expand_as(tran_feat)
        feat = feat.view(B, N, C, fH, fW).permute(0, 1, 3, 4, 2).contiguous()
        dz, dy, dx = int(size[2]), int(size[1]), int(size[0])
        g = out_grads[p]
        if collapse_z:                     # (B, Dz*C, Dy, Dx) -> (B, Dz, Dy, Dx, C): undo cat(unbind(2), 1) + permute
            g = g.reshape(B, dz, C, dy, dx).permute(0, 1, 3, 4, 2)
        else:                              # (B, C, Dz, Dy, Dx) -> (B, Dz, Dy, Dx, C)
            g = g.permute(0, 2, 3, 4, 1)
        dg, fg = bev_pool_v2_grad_ref(g, d5, feat, rd, rf, rb)
        dgrad += dg
        fg = fg.view(B * N, fH, fW, C).permute(0, 3, 1, 2)
        fgrad += fg if p == 0 else fg * masks[p - 1].unsqueeze(1)
    return dgrad.view(B * N, D, fH, fW), fgrad
