"""The hot path as one re-runnable step (used by bench.py and the multi-GPU harness).

Stages implemented here run only through the C-ABI library; torch provides device
memory, pinned host buffers, streams and events.  Per step and per GPU (DHD-S, B samples):

  height_to_mask   argmax over the height distribution -> per-pixel mask id      (a5)
  mghs_prepare     fused get_ego_coor + quantise + bin by BEV cell, once for all
                   four passes (geom_count, scan_local, scan_blocks, scatter,
                   canonical)                                                     (a7, a8)
  mghs_pool_fwd    masked lift (x) splat, four grids, single write              (a6, a9, a11)
  mghs_pool_bwd    depth / context gradients of the four passes                   (a10)
"""
import json
import os

import torch

from .pool import MghsPool, height_to_mask

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def algorithmic_bytes(cfg, B):
    """SURVEY.md 8(d) / BASELINE.md section 3: every output element written exactly once,
    every input element read once; index tables, bins and workspaces are not counted."""
    N = cfg['ncams']
    h, w = cfg['input_size']
    fH, fW = h // cfg['downsample'], w // cfg['downsample']
    D = len(range(int(cfg['depth'][0]), int(cfg['depth'][1]), int(cfg['depth'][2])))
    C = cfg['C']
    P = B * N * fH * fW
    grids = [cfg['bev_grid']] + list(cfg['mask_grids'])
    out_elems = 0
    for g in grids:
        dx = int(round((g['x'][1] - g['x'][0]) / g['x'][2]))
        dy = int(round((g['y'][1] - g['y'][0]) / g['y'][2]))
        dz = int(round((g['z'][1] - g['z'][0]) / g['z'][2]))
        out_elems += B * dz * dy * dx * C
    reads = P * D * 4 + P * C * 4 + P
    return {'pool_fwd_bytes': reads + out_elems * 4,
            'pool_bwd_bytes': out_elems * 4 + reads + P * D * 4 + P * C * 4,
            'pool_out_bytes': out_elems * 4}


class HotPathStep:
    def __init__(self, cfg, B, layout='nhwc', deterministic=True, device='cuda'):
        self.cfg, self.B, self.layout, self.deterministic = cfg, B, layout, deterministic
        self.N = cfg['ncams']
        h, w = cfg['input_size']
        self.fH, self.fW = h // cfg['downsample'], w // cfg['downsample']
        self.D = torch.arange(*cfg['depth']).numel()
        self.C = cfg['C']
        grids = [cfg['bev_grid']] + list(cfg['mask_grids'])
        self.plan = MghsPool(B, self.N, self.D, self.fH, self.fW, self.C, grids[0]['x'], grids[0]['y'],
                             [(g['z'], m) for m, g in enumerate(grids)])
        self.device = torch.device(device)
        self.workspace = torch.empty(self.plan.ws_bytes, dtype=torch.uint8, device=self.device)
        self.outs = self.plan.alloc_outputs(layout, self.device)
        g = torch.Generator(device=self.device).manual_seed(7)
        # upstream gradients of the four BEV tensors (what the encoders hand back), resident
        self.gouts = [torch.randn(o.shape if layout == 'nhwc' else
                                  (B, self.plan.Dy, self.plan.Dx, dz * self.C),
                                  device=self.device, generator=g)
                      for o, dz in zip(self.outs, self.plan.dz)]
        self.depth_grad = torch.empty(B * self.N, self.D, self.fH, self.fW, device=self.device)
        self.feat_grad = torch.empty(B, self.N, self.fH, self.fW, self.C, device=self.device)
        self.host_dgrad = torch.empty(self.depth_grad.shape, pin_memory=True)
        self.host_fgrad = torch.empty(self.feat_grad.shape, pin_memory=True)
        self.height_range = cfg['height_range']
        self.mask_range = cfg['mask_range']
        self.h2d_bytes = 0
        self.d2h_bytes = self.host_dgrad.numel() * 4 + self.host_fgrad.numel() * 4
        self.launches_per_step = 1 + (5 if deterministic else 4) + 1 + 1
        self._dev_e2e = None

    def stage_names(self):
        return ['height_to_mask', 'mghs_prepare(geometry+binning, all 4 passes)',
                'mghs_pool_fwd(%s)' % self.layout, 'mghs_pool_bwd(nhwc grads)']

    # ---- inputs ---------------------------------------------------------------------
    def pin_host_inputs(self, inputs, depth, feat, height):
        """Host-side (pinned) step inputs in the layouts the plugin call takes."""
        from .pool import grid_infos  # noqa: F401  (kept for symmetry with MGHS)
        x, s2e, e2g, K, pr, pt, bda = inputs
        B, N, C = self.B, self.N, self.C
        h = {
            'depth': depth.contiguous(),
            'feat': feat.view(B, N, C, self.fH, self.fW).permute(0, 1, 3, 4, 2).contiguous(),
            'height': height.contiguous(),
            'sensor2ego': s2e.contiguous(), 'cam2imgs': K.contiguous(), 'post_rots': pr.contiguous(),
            'post_trans': pt.contiguous(), 'bda': bda.contiguous(),
        }
        h = {k: v.pin_memory() for k, v in h.items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in h.values())
        return h

    def frustum(self):
        """create_frustum (lss_heightmap.py:105-134), sid=False."""
        cfg = self.cfg
        h_in, w_in = cfg['input_size']
        d = torch.arange(*cfg['depth'], dtype=torch.float).view(-1, 1, 1).expand(-1, self.fH, self.fW)
        xs = torch.linspace(0, w_in - 1, self.fW, dtype=torch.float).view(1, 1, self.fW).expand(self.D, self.fH, self.fW)
        ys = torch.linspace(0, h_in - 1, self.fH, dtype=torch.float).view(1, self.fH, 1).expand(self.D, self.fH, self.fW)
        return torch.stack((xs, ys, d), -1)

    def to_device(self, host):
        dev = {k: v.to(self.device, non_blocking=True) for k, v in host.items()}
        dev['frustum'] = self.frustum().to(self.device)
        return dev

    # ---- one step ----------------------------------------------------------------------
    def run(self, dev, pool_events=None):
        pixmask = height_to_mask(dev['height'], self.height_range, self.mask_range)
        self.plan.prepare(frustum=dev['frustum'], sensor2ego=dev['sensor2ego'],
                          cam2imgs=dev['cam2imgs'], post_rots=dev['post_rots'],
                          post_trans=dev['post_trans'], bda=dev['bda'],
                          deterministic=self.deterministic, workspace=self.workspace)
        st = torch.cuda.current_stream()
        if pool_events is not None:
            pool_events[0].record(st)
        self.plan.raw_forward(dev['depth'], dev['feat'], pixmask, self.outs, self.layout)
        if pool_events is not None:
            pool_events[1].record(st)
        import ctypes
        from . import _lib
        arr = (ctypes.c_void_p * len(self.gouts))(*[g.data_ptr() for g in self.gouts])
        _lib.check(self.plan._lib.dhd_mghs_pool_bwd(
            ctypes.byref(self.plan.cfg), ctypes.c_void_p(dev['depth'].data_ptr()),
            ctypes.c_void_p(dev['feat'].data_ptr()), ctypes.c_void_p(pixmask.data_ptr()),
            ctypes.c_void_p(self.workspace.data_ptr()), arr, 0,
            ctypes.c_void_p(self.depth_grad.data_ptr()), ctypes.c_void_p(self.feat_grad.data_ptr()),
            ctypes.c_void_p(st.cuda_stream)), 'mghs_pool_bwd')
        self._keep = pixmask

    def run_e2e(self, host):
        """Host buffers in, host buffers out: H2D of every step input, the step, D2H of the
        step result (depth / context gradients)."""
        if self._dev_e2e is None:
            self._dev_e2e = {'frustum': self.frustum().to(self.device)}
        d = self._dev_e2e
        for k, v in host.items():
            if k not in d:
                d[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
            d[k].copy_(v, non_blocking=True)
        self.run(d)
        self.host_dgrad.copy_(self.depth_grad, non_blocking=True)
        self.host_fgrad.copy_(self.feat_grad, non_blocking=True)

    def ncu_traffic_bytes(self):
        """dram read+write bytes per launch of the pool kernel from the committed ncu capture."""
        p = os.path.join(_ROOT, 'profiles', 'pool_fwd_%s_traffic.json' % self.layout)
        if os.path.exists(p):
            try:
                return json.load(open(p)).get('dram_bytes_per_launch')
            except Exception:
                return None
        return None
