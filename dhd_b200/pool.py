"""Host-side mirror of the reference operator interface for the voxel-pool hot path.

``bev_pool_v2`` / ``QuickCumsumCuda`` keep the reference names, argument meaning and
error behaviour (projects/mmdet3d_plugin/ops/bev_pool_v2/bev_pool.py:11-106);
``MghsPool`` is the fused replacement of the four-pass ``MGHS.view_transform`` body
(models/necks/lss_heightmap.py:261-371, 407-459).  Both call the C-ABI library only:
torch is used for device memory, streams and autograd bookkeeping.
"""
import ctypes

import torch

from . import _lib
from ._lib import LAYOUT_NCDHW, LAYOUT_NCDHW_CAT, LAYOUT_NCHW_COLLAPSE, LAYOUT_NHWC, LAYOUT_NHWC_BF16, MghsCfg

_LAYOUTS = {'nhwc': LAYOUT_NHWC, 'nchw': LAYOUT_NCHW_COLLAPSE, 'ncdhw': LAYOUT_NCDHW, 'ncdhw_cat': LAYOUT_NCDHW_CAT,
            'nhwc_bf16': LAYOUT_NHWC_BF16}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback); '
                               'got a %s tensor' % t.device)


# ------------------------------------------------------------------ drop-in operator
class QuickCumsumCuda(torch.autograd.Function):
    """BEVPoolv2 (reference: ops/bev_pool_v2/bev_pool.py:11-83)."""

    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                interval_starts, interval_lengths):
        _need_cuda(depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths)
        ranks_bev = ranks_bev.int().contiguous()
        depth = depth.contiguous().float()
        feat = feat.contiguous().float()
        ranks_depth = ranks_depth.contiguous().int()
        ranks_feat = ranks_feat.contiguous().int()
        interval_lengths = interval_lengths.contiguous().int()
        interval_starts = interval_starts.contiguous().int()
        out = feat.new_zeros(bev_feat_shape)
        lib = _lib.load()
        _lib.check(lib.dhd_bev_pool_v2_fwd(
            feat.shape[-1], interval_starts.numel(), _ptr(depth), _ptr(feat), _ptr(ranks_depth),
            _ptr(ranks_feat), _ptr(ranks_bev), _ptr(interval_starts), _ptr(interval_lengths),
            _ptr(out), _stream()), 'bev_pool_v2_fwd')
        ctx.save_for_backward(ranks_bev, depth, feat, ranks_feat, ranks_depth)
        return out

    @staticmethod
    def backward(ctx, out_grad):
        ranks_bev, depth, feat, ranks_feat, ranks_depth = ctx.saved_tensors
        # same re-grouping by feature index as the reference (bev_pool.py:47-57)
        order = ranks_feat.argsort()
        ranks_feat, ranks_depth, ranks_bev = ranks_feat[order], ranks_depth[order], ranks_bev[order]
        kept = torch.ones(ranks_bev.shape[0], device=ranks_bev.device, dtype=torch.bool)
        kept[1:] = ranks_feat[1:] != ranks_feat[:-1]
        starts = torch.where(kept)[0].int()
        lengths = torch.zeros_like(starts)
        lengths[:-1] = starts[1:] - starts[:-1]
        lengths[-1] = ranks_bev.shape[0] - starts[-1]
        depth_grad = depth.new_zeros(depth.shape)
        feat_grad = feat.new_zeros(feat.shape)
        out_grad = out_grad.contiguous()
        lib = _lib.load()
        _lib.check(lib.dhd_bev_pool_v2_bwd(
            feat.shape[-1], starts.numel(), _ptr(out_grad), _ptr(depth), _ptr(feat),
            _ptr(ranks_depth.contiguous()), _ptr(ranks_feat.contiguous()),
            _ptr(ranks_bev.contiguous()), _ptr(starts.contiguous()), _ptr(lengths.contiguous()),
            _ptr(depth_grad), _ptr(feat_grad), _stream()), 'bev_pool_v2_bwd')
        return depth_grad, feat_grad, None, None, None, None, None, None


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                interval_starts, interval_lengths):
    """Same contract as the reference (bev_pool.py:86-106).

    depth (B,N,D,fH,fW); feat (B,N,fH,fW,C); ranks_* (N_points,) ; bev_feat_shape
    (B,Dz,Dy,Dx,C); interval_* (N_pillar,).  Returns (B, C, Dz, Dy, Dx) contiguous fp32."""
    x = QuickCumsumCuda.apply(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                              interval_starts, interval_lengths)
    return x.permute(0, 4, 1, 2, 3).contiguous()


# ------------------------------------------------------------------ height -> mask id
_TABLES = {}


def height_to_mask(height, height_range, thresholds):
    """(BN, H, fH, fW) height distribution -> int8 (BN, fH, fW) mask id: k (1-based) iff
    thresholds[k-1] <= height_range[argmax] < thresholds[k], else 0
    (lss_heightmap.py:528-564; comparisons in fp32 as torch does for float32-vs-scalar)."""
    _need_cuda(height)
    height = height.contiguous().float()
    BN, H, fH, fW = height.shape
    key = (tuple(height_range), tuple(thresholds), height.device)
    if key not in _TABLES:       # cached: no pageable H2D copy per call (CUDA-graph capturable)
        _TABLES[key] = (torch.tensor(height_range, dtype=torch.float32, device=height.device),
                        torch.tensor(thresholds, dtype=torch.float32, device=height.device))
    hr, th = _TABLES[key]
    if hr.numel() != H:
        raise ValueError('height_range has %d entries, height has %d channels' % (hr.numel(), H))
    out = torch.empty(BN, fH, fW, dtype=torch.int8, device=height.device)
    _lib.check(_lib.load().dhd_height_to_mask(_ptr(height), BN, H, fH * fW, _ptr(hr), _ptr(th),
                                              th.numel() - 1, _ptr(out), _stream()),
               'height_to_mask')
    return out


# ------------------------------------------------------------------------- fused pool
def grid_infos(axis_cfgs):
    """lower / interval / size exactly as create_grid_infos builds them: Python-float
    arithmetic rounded to fp32 by torch.Tensor(...) (lss_heightmap.py:99-102)."""
    lower = torch.Tensor([c[0] for c in axis_cfgs])
    interval = torch.Tensor([c[2] for c in axis_cfgs])
    size = torch.Tensor([(c[1] - c[0]) / c[2] for c in axis_cfgs])
    return lower, interval, size


class _MghsPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, feat, plan, pixmask, layout):
        outs = plan._forward(depth, feat, pixmask, layout)
        ctx.plan, ctx.pixmask, ctx.layout = plan, pixmask, layout
        ctx.workspace = plan.workspace      # keep the bins this forward used alive
        ctx.save_for_backward(depth, feat)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        depth, feat = ctx.saved_tensors
        dg, fg = ctx.plan._backward(depth, feat, ctx.pixmask, gouts, ctx.layout, ctx.workspace)
        return dg, fg, None, None, None


class MghsPool:
    """Fused multi-pass Lift-Splat pool.

    passes: list of (z_cfg=[lower, upper, step], mask_id) -- mask_id 0 pools every pixel,
    k>0 only pixels whose height mask is k.  All passes share the x / y grid.
    """

    def __init__(self, B, N, D, fH, fW, C, x_cfg, y_cfg, passes):
        if not 1 <= len(passes) <= _lib.MAX_PASSES:
            raise ValueError('1..%d passes supported' % _lib.MAX_PASSES)
        self.B, self.N, self.D, self.fH, self.fW, self.C = B, N, D, fH, fW, C
        self.passes = [(list(z), int(m)) for z, m in passes]
        cfg = MghsCfg()
        cfg.B, cfg.N, cfg.D, cfg.fH, cfg.fW, cfg.C = B, N, D, fH, fW, C
        lo, iv, sz = grid_infos([x_cfg, y_cfg])
        cfg.x_lower, cfg.x_interval, cfg.x_size = float(lo[0]), float(iv[0]), float(sz[0])
        cfg.y_lower, cfg.y_interval, cfg.y_size = float(lo[1]), float(iv[1]), float(sz[1])
        cfg.Dx, cfg.Dy = int(sz[0]), int(sz[1])
        cfg.n_pass = len(passes)
        self.dz = []
        for p, (z, m) in enumerate(self.passes):
            zl, zi, zs = grid_infos([z])
            cfg.z_lower[p], cfg.z_interval[p], cfg.z_size[p] = float(zl[0]), float(zi[0]), float(zs[0])
            cfg.dz[p] = int(zs[0])
            cfg.mask_id[p] = m
            self.dz.append(int(zs[0]))
        self.Dx, self.Dy = cfg.Dx, cfg.Dy
        self.cfg = cfg
        self.masked = any(m != 0 for _, m in self.passes)
        self.workspace = None
        self._lib = _lib.load()
        self.ws_bytes = self._lib.dhd_mghs_workspace_bytes(ctypes.byref(cfg))
        if self.ws_bytes == 0:
            raise RuntimeError('dhd_b200.mghs: ' + self._lib.dhd_last_error().decode())
        self.npoints = B * N * D * fH * fW

    # -- binning ---------------------------------------------------------------------
    @staticmethod
    def camera_matrices(sensor2ego, cam2imgs, post_rots, post_trans, bda):
        """Per-camera 3x3s with the reference's own torch calls (lss_heightmap.py:207, 217):
        inv(post_rots), post_trans, sensor2ego[:3,:3] @ inv(K), sensor2ego[:3,3], bda.
        linalg.inv_ex is torch.inverse without the error-check host sync (same bits)."""
        B, N = sensor2ego.shape[:2]
        ipr = torch.linalg.inv_ex(post_rots).inverse.reshape(B * N, 3, 3)
        comb = sensor2ego[:, :, :3, :3].matmul(torch.linalg.inv_ex(cam2imgs).inverse)
        return (ipr.contiguous().float(), post_trans.reshape(B * N, 3).contiguous().float(),
                comb.reshape(B * N, 3, 3).contiguous().float(),
                sensor2ego[:, :, :3, 3].reshape(B * N, 3).contiguous().float(),
                bda.contiguous().float())

    def prepare(self, frustum=None, sensor2ego=None, cam2imgs=None, post_rots=None,
                post_trans=None, bda=None, coor=None, cam_mats=None, deterministic=True,
                workspace=None):
        """Bin the frustum by BEV cell.  Geometry source, in order of precedence:
        `coor` (B,N,D,fH,fW,3), the result of the reference's get_ego_coor; `cam_mats`, the
        tuple camera_matrices() returns; or the raw camera tensors.  With the last two the
        per-point transform runs fused in the kernel."""
        src = coor if coor is not None else (cam_mats[0] if cam_mats is not None else sensor2ego)
        dev = src.device
        if dev.type != 'cuda':
            raise RuntimeError('dhd_b200.mghs: CUDA tensors required (no CPU fallback)')
        ws = workspace
        if ws is None:
            ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        args = [None] * 9
        if coor is not None:
            coor = coor.contiguous().float()
            if coor.numel() != self.npoints * 3:
                raise ValueError('coor has the wrong number of points')
            args[0] = coor
        else:
            if cam_mats is None:
                cam_mats = self.camera_matrices(sensor2ego, cam2imgs, post_rots, post_trans, bda)
            ipr, ptn, comb, tr, bda3 = [m.contiguous().float() for m in cam_mats]
            if ipr.numel() != self.B * self.N * 9 or bda3.numel() != self.B * 9:
                raise ValueError('camera matrices do not match (B, N)')
            fkey = (frustum.data_ptr(), frustum._version, str(dev))
            if getattr(self, '_fvec_key', None) != fkey:     # the frustum is a constant buffer: slice it once
                f = frustum.to(device=dev, dtype=torch.float32)
                self._fvec = (f[0, 0, :, 0].contiguous(), f[0, :, 0, 1].contiguous(), f[:, 0, 0, 2].contiguous())
                self._fvec_key = fkey
            fu, fv, fd = self._fvec
            args[1:] = [fu, fv, fd, ipr, ptn, comb, tr, bda3]
        _lib.check(self._lib.dhd_mghs_prepare(
            ctypes.byref(self.cfg), *[_ptr(a) for a in args], _ptr(ws), int(bool(deterministic)),
            _stream()), 'mghs_prepare')
        self.workspace = ws
        self._keep = args       # inputs must outlive the enqueued kernels
        return ws

    def num_entries(self):
        off = self._lib.dhd_mghs_workspace_count_offset(ctypes.byref(self.cfg))
        return int(self.workspace[off:off + 4].view(torch.int32).item())

    def voxel_index(self):
        """(n_pass, F) int32 voxel rank of every frustum point per pass, -1 = not kept."""
        out = torch.empty(len(self.passes), self.npoints, dtype=torch.int32,
                          device=self.workspace.device)
        _lib.check(self._lib.dhd_mghs_voxel_index(ctypes.byref(self.cfg), _ptr(self.workspace),
                                                  _ptr(out), _stream()), 'mghs_voxel_index')
        return out

    # -- pooling -----------------------------------------------------------------------
    def alloc_outputs(self, layout, device):
        if layout == 'ncdhw_cat':     # pass 0 alone, passes 1.. stacked on z (MGHS_Depth, LH:845)
            return [torch.empty(self.B, self.C, self.dz[0], self.Dy, self.Dx, device=device),
                    torch.empty(self.B, self.C, sum(self.dz[1:]), self.Dy, self.Dx, device=device)]
        outs = []
        for dz in self.dz:
            if layout == 'nhwc':
                outs.append(torch.empty(self.B, self.Dy, self.Dx, dz * self.C, device=device))
            elif layout == 'nhwc_bf16':        # forward only: the encoders' input activations
                outs.append(torch.empty(self.B, self.Dy, self.Dx, dz * self.C, device=device, dtype=torch.bfloat16))
            elif layout == 'nchw':
                outs.append(torch.empty(self.B, dz * self.C, self.Dy, self.Dx, device=device))
            elif layout == 'ncdhw':
                outs.append(torch.empty(self.B, self.C, dz, self.Dy, self.Dx, device=device))
            else:
                raise ValueError('layout must be nhwc, nchw or ncdhw')
        return outs

    def _check_inputs(self, depth, feat, pixmask):
        _need_cuda(depth, feat, pixmask)
        if self.workspace is None:
            raise RuntimeError('dhd_b200.mghs: call prepare() before pooling')
        if depth.numel() != self.npoints:
            raise ValueError('depth must have B*N*D*fH*fW = %d elements' % self.npoints)
        if feat.numel() != self.npoints // self.D * self.C or feat.shape[-1] != self.C:
            raise ValueError('feat must be (B, N, fH, fW, C) with C=%d' % self.C)
        if self.masked and pixmask is None:
            raise ValueError('a pass is masked: pixmask is required')
        if pixmask is not None and (pixmask.dtype != torch.int8 or
                                    pixmask.numel() != self.npoints // self.D):
            raise ValueError('pixmask must be int8 with B*N*fH*fW elements')

    def raw_forward(self, depth, feat, pixmask, outs, layout='nhwc', workspace=None):
        """Enqueue the pool kernel into preallocated outputs (no autograd, no allocation)."""
        ptrs = [o.data_ptr() for o in outs]
        ptrs += [0] * (len(self.passes) - len(ptrs))          # 'ncdhw_cat' hands over two tensors
        arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        _lib.check(self._lib.dhd_mghs_pool_fwd(
            ctypes.byref(self.cfg), _ptr(depth), _ptr(feat), _ptr(pixmask),
            _ptr(workspace if workspace is not None else self.workspace), arr, _LAYOUTS[layout],
            _stream()), 'mghs_pool_fwd')

    def _forward(self, depth, feat, pixmask, layout):
        depth = depth.contiguous().float()
        feat = feat.contiguous().float()
        if pixmask is not None:
            pixmask = pixmask.contiguous()
        self._check_inputs(depth, feat, pixmask)
        outs = self.alloc_outputs(layout, depth.device)
        self.raw_forward(depth, feat, pixmask, outs, layout)
        return outs

    def _backward(self, depth, feat, pixmask, gouts, layout, workspace):
        if layout == 'ncdhw_cat':     # split the stacked gradient back into per-pass (B, C, dz, Dy, Dx) views
            g0, gc = gouts
            parts, z = [], 0
            for dz in self.dz[1:]:
                parts.append(None if gc is None else gc[:, :, z:z + dz])
                z += dz
            gouts, layout = [g0] + parts, 'ncdhw'
        gs = []
        for g, dz in zip(gouts, self.dz):
            if g is None:
                g = torch.zeros(self.B, self.Dy, self.Dx, dz * self.C, device=depth.device)
            elif layout == 'nhwc':
                g = g.contiguous().float()
            elif layout == 'nchw':        # (B, dz*C, Dy, Dx) -> (B, Dy, Dx, dz*C)
                g = g.float().permute(0, 2, 3, 1).contiguous()
            else:                         # (B, C, dz, Dy, Dx) -> (B, Dy, Dx, dz, C)
                g = g.float().permute(0, 3, 4, 2, 1).contiguous()
            gs.append(g)
        depth = depth.contiguous().float()
        feat = feat.contiguous().float()
        dg = torch.empty_like(depth)
        fg = torch.empty_like(feat)
        arr = (ctypes.c_void_p * len(gs))(*[g.data_ptr() for g in gs])
        _lib.check(self._lib.dhd_mghs_pool_bwd(
            ctypes.byref(self.cfg), _ptr(depth), _ptr(feat), _ptr(pixmask), _ptr(workspace), arr,
            LAYOUT_NHWC, _ptr(dg), _ptr(fg), _stream()), 'mghs_pool_bwd')
        return dg, fg

    def __call__(self, depth, feat, pixmask=None, layout='nhwc'):
        """depth (B,N,D,fH,fW) or (B*N,D,fH,fW); feat (B,N,fH,fW,C) channels-last context;
        pixmask int8 (B*N,fH,fW).  Returns one tensor per pass in the memory `layout`:
        'nhwc' (B,Dy,Dx,dz*C) | 'nchw' (B,dz*C,Dy,Dx) | 'ncdhw' (B,C,dz,Dy,Dx) | 'ncdhw_cat' (two
        tensors: pass 0 as ncdhw, passes 1.. stacked on z).
        Differentiable w.r.t. depth and feat."""
        return _MghsPoolFn.apply(depth, feat, self, pixmask, layout)
