"""MGHS view transformer: reference-compatible constructor, attributes, parameter names and
return values (projects/mmdet3d_plugin/models/necks/lss_heightmap.py:12-701), with the forward
pass re-designed for B200:

  reference forward (LH:461-490, 407-459)            this module
  ------------------------------------------------   -------------------------------------------
  depth_net conv, slice, softmax                     ONE tcgen05 GEMM, softmax + split in epilogue
  HeightNet (14 cuDNN launches) + softmax            tcgen05 conv chain, softmax in last epilogue
  argmax -> height map -> 3 masks -> 3 masked copies dhd_height_to_mask: one int8 mask id per pixel
  4 x (get_ego_coor + quantise + argsort + zeros +   dhd_mghs_prepare (geometry + binning, once for
       bev_pool_v2 + permute + collapse-Z cat)       the four grids) + dhd_mghs_pool_fwd (one fused,
                                                     single-write kernel)

The reference-shaped single-pass methods (get_ego_coor, voxel_pooling_prepare_v2,
voxel_pooling_v2, view_transform_core) are kept for API parity and run on the drop-in
``bev_pool_v2`` operator.
"""
import torch
import torch.nn as nn

from dhd_b200.compat import NECKS, BaseModule, EngineOwner, force_fp32
from dhd_b200.pool import MghsPool, height_to_mask

from ...ops import bev_pool_v2
from ..model_utils import DepthNet, HeightNet

_BEV_PASS_GRID = {'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [-1, 5.4, 6.4],
                  'depth': [1.0, 45.0, 0.5]}      # hard-coded by the reference, LH:425-431


@NECKS.register_module(force=True)
class MGHS(EngineOwner, BaseModule):
    def __init__(self, grid_config, input_size, downsample=16, in_channels=512, out_channels=64,
                 heightnet_cfg=dict(), accelerate=False, sid=False, collapse_z=True,
                 height_range=[-1.5, -1, 0, 0.5, 1, 1.5, 2, 2.5, 3, 3.5, 4], height_interval=0.5,
                 mask_range=[-5, 0, 0.4, 5], loss_height_weight=1.0,
                 mask_1_grid={'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [-1, 2.2, 0.4], 'depth': [1.0, 45.0, 0.5]},
                 mask_2_grid={'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [2.2, 3.8, 0.4], 'depth': [1.0, 45.0, 0.5]},
                 mask_3_grid={'x': [-40, 40, 0.4], 'y': [-40, 40, 0.4], 'z': [3.8, 5.4, 0.4], 'depth': [1.0, 45.0, 0.5]},
                 precision='fp32', out_layout='nhwc'):
        super().__init__()
        self.grid_config = grid_config
        self.downsample = downsample
        self.input_size = input_size
        self.create_grid_infos(**grid_config)
        self.sid = sid
        self.frustum = self.create_frustum(grid_config['depth'], input_size, downsample)
        self.accelerate = accelerate
        self.initial_flag = True
        self.out_channels, self.in_channels = out_channels, in_channels
        self.depth_net = nn.Conv2d(in_channels, self.D + out_channels, kernel_size=1, padding=0)
        self.H = len(height_range)
        self.precision = precision
        self.height_net = HeightNet(in_channels=in_channels, mid_channels=in_channels,
                                    depth_channels=self.H, precision=precision, **heightnet_cfg)
        self.collapse_z = collapse_z
        self.height_range, self.mask_range = height_range, mask_range
        self.height_interval = height_interval
        self.loss_height_weight = loss_height_weight
        self.mask_1_grid, self.mask_2_grid, self.mask_3_grid = mask_1_grid, mask_2_grid, mask_3_grid
        self.out_layout = out_layout          # 'nhwc' (channels_last memory) or 'nchw' (reference memory)
        self._plan = None
        self._depth_engine = None

    # ------------------------------------------------------------------ grids / frustum (LH:86-134)
    def create_grid_infos(self, x, y, z, **kwargs):
        self.grid_lower_bound = torch.Tensor([c[0] for c in (x, y, z)])
        self.grid_interval = torch.Tensor([c[2] for c in (x, y, z)])
        self.grid_size = torch.Tensor([(c[1] - c[0]) / c[2] for c in (x, y, z)])

    def create_frustum(self, depth_cfg, input_size, downsample):
        h_in, w_in = input_size
        fh, fw = h_in // downsample, w_in // downsample
        d = torch.arange(*depth_cfg, dtype=torch.float)
        self.D = d.shape[0]
        if self.sid:
            lo, hi = float(depth_cfg[0]), float(depth_cfg[1])
            k = torch.arange(self.D).float()
            d = torch.exp(torch.log(torch.tensor(lo)) + k / (self.D - 1) * torch.log(torch.tensor((hi - 1) / lo)))
        u = torch.linspace(0, w_in - 1, fw, dtype=torch.float)
        v = torch.linspace(0, h_in - 1, fh, dtype=torch.float)
        fr = torch.empty(self.D, fh, fw, 3)
        fr[..., 0], fr[..., 1], fr[..., 2] = u.view(1, 1, fw), v.view(1, fh, 1), d.view(-1, 1, 1)
        return fr

    # ------------------------------------------------------------------ reference-shaped single pass
    def get_ego_coor(self, sensor2ego, ego2global, cam2imgs, post_rots, post_trans, bda):
        """(B, N, D, fH, fW, 3) ego-frame frustum points; same torch ops as LH:179-231."""
        B, N = sensor2ego.shape[:2]
        p = self.frustum.to(sensor2ego) - post_trans.view(B, N, 1, 1, 1, 3)
        p = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(p.unsqueeze(-1))
        p = torch.cat((p[..., :2, :] * p[..., 2:3, :], p[..., 2:3, :]), 5)
        comb = sensor2ego[:, :, :3, :3].matmul(torch.inverse(cam2imgs))
        p = comb.view(B, N, 1, 1, 1, 3, 3).matmul(p).squeeze(-1)
        p = p + sensor2ego[:, :, :3, 3].view(B, N, 1, 1, 1, 3)
        return bda.view(B, 1, 1, 1, 1, 3, 3).matmul(p.unsqueeze(-1)).squeeze(-1)

    get_lidar_coor = get_ego_coor

    def voxel_pooling_prepare_v2(self, coor):
        """ranks + run-length intervals of one grid (LH:303-371): fp32 subtract-then-divide,
        truncation toward zero, fp32 rank arithmetic, argsort."""
        B, N, D, H, W, _ = coor.shape
        n = B * N * D * H * W
        dev = coor.device
        ranks_depth = torch.arange(n, dtype=torch.int, device=dev)
        ranks_feat = torch.arange(n // D, dtype=torch.int, device=dev).reshape(B, N, 1, H, W)
        ranks_feat = ranks_feat.expand(B, N, D, H, W).flatten()
        idx = ((coor - self.grid_lower_bound.to(coor)) / self.grid_interval.to(coor)).long().view(n, 3)
        batch = torch.arange(B, device=dev).view(B, 1).expand(B, n // B).reshape(n, 1)
        size = self.grid_size.to(coor)
        kept = ((idx[:, 0] >= 0) & (idx[:, 0] < size[0]) & (idx[:, 1] >= 0) & (idx[:, 1] < size[1]) &
                (idx[:, 2] >= 0) & (idx[:, 2] < size[2]))
        if int(kept.sum()) == 0:
            return None, None, None, None, None
        idx, batch = idx[kept], batch[kept, 0]
        ranks_depth, ranks_feat = ranks_depth[kept], ranks_feat[kept]
        rb = batch * (size[2] * size[1] * size[0]) + idx[:, 2] * (size[1] * size[0]) + \
            idx[:, 1] * size[0] + idx[:, 0]
        order = rb.argsort()
        rb, ranks_depth, ranks_feat = rb[order], ranks_depth[order], ranks_feat[order]
        first = torch.ones(rb.shape[0], device=dev, dtype=torch.bool)
        first[1:] = rb[1:] != rb[:-1]
        starts = torch.where(first)[0].int()
        lengths = torch.zeros_like(starts)
        lengths[:-1] = starts[1:] - starts[:-1]
        lengths[-1] = rb.shape[0] - starts[-1]
        return (rb.int().contiguous(), ranks_depth.int().contiguous(), ranks_feat.int().contiguous(),
                starts.int().contiguous(), lengths.int().contiguous())

    def voxel_pooling_v2(self, coor, depth, feat):
        """One grid through the drop-in operator (LH:261-300)."""
        rb, rd, rf, st, ln = self.voxel_pooling_prepare_v2(coor)
        dz, dy, dx = (int(self.grid_size[i]) for i in (2, 1, 0))
        if rf is None:
            print('warning ---> no points within the predefined bev receptive field')
            out = torch.zeros(feat.shape[0], feat.shape[2], dz, dy, dx).to(feat)
        else:
            out = bev_pool_v2(depth, feat.permute(0, 1, 3, 4, 2), rd, rf, rb,
                              (depth.shape[0], dz, dy, dx, feat.shape[2]), st, ln)
        return torch.cat(out.unbind(dim=2), 1) if self.collapse_z else out

    def view_transform_core(self, input, depth, tran_feat):
        B, N, C, H, W = input[0].shape
        coor = self.get_ego_coor(*input[1:7])
        bev = self.voxel_pooling_v2(coor, depth.view(B, N, self.D, H, W),
                                    tran_feat.view(B, N, self.out_channels, H, W))
        return bev, depth

    def init_acceleration_v2(self, coor):
        """LH:234-258: the rank / interval tensors of one grid kept as attributes (what the reference's, unused,
        accelerate path would read).  The fused path caches its bins through `pre_compute` instead."""
        ranks_bev, ranks_depth, ranks_feat, starts, lengths = self.voxel_pooling_prepare_v2(coor)
        self.ranks_bev = ranks_bev.int().contiguous()
        self.ranks_feat = ranks_feat.int().contiguous()
        self.ranks_depth = ranks_depth.int().contiguous()
        self.interval_starts = starts.int().contiguous()
        self.interval_lengths = lengths.int().contiguous()

    def pre_compute(self, input):
        """Bins depend on the camera geometry only: cache them (the reference's `accelerate`)."""
        if self.initial_flag:
            self._prepare(input)
            self.initial_flag = not self.accelerate     # without accelerate every forward re-bins anyway

    # ------------------------------------------------------------------ height masks (LH:528-564)
    def height_feature_to_height_map(self, height_feature, height_range):
        if height_feature.dim() != 4:
            raise ValueError('Input tensor must have 4 dimensions (BxN, H, fH, fW)')
        return torch.tensor(height_range, device=height_feature.device)[torch.argmax(height_feature, dim=1)]

    def create_mask_3(self, input_tensor, h_min, thr1, thr2, h_max):
        return ((input_tensor >= h_min) & (input_tensor < thr1), (input_tensor >= thr1) & (input_tensor < thr2),
                (input_tensor >= thr2) & (input_tensor < h_max))

    # ------------------------------------------------------------------ fused four-pass transform
    def _passes(self):
        grids = [_BEV_PASS_GRID, self.mask_1_grid, self.mask_2_grid, self.mask_3_grid]
        return grids, [(g['z'], m) for m, g in enumerate(grids)]

    def _get_plan(self, B, N, H, W):
        key = (B, N, H, W)
        if self._plan is None or self._plan[0] != key:
            grids, passes = self._passes()
            for g in grids[1:]:
                if list(g['x']) != list(grids[0]['x']) or list(g['y']) != list(grids[0]['y']):
                    raise NotImplementedError('MGHS fused pool: the mask grids must share the BEV x/y grid')
            self._plan = (key, MghsPool(B, N, self.D, H, W, self.out_channels, grids[0]['x'], grids[0]['y'], passes))
        return self._plan[1]

    def _prepare(self, input):
        B, N, _, H, W = input[0].shape
        plan = self._get_plan(B, N, H, W)
        s2e, _e2g, K, pr, pt, bda = input[1:7]
        plan.prepare(frustum=self.frustum, sensor2ego=s2e.float(), cam2imgs=K.float(),
                     post_rots=pr.float(), post_trans=pt.float(), bda=bda.float())
        return plan

    def _bins(self, input):
        """The binned frustum of this call.  accelerate=True (the reference's dead flag, LH:56, 374-378): the bins of
        the first call -- or of pre_compute() -- are reused while (B, N, fH, fW) stays the same; a different shape
        (e.g. a smaller last batch) re-bins.  The camera geometry is assumed fixed, as the reference's flag does."""
        B, N, _, H, W = input[0].shape
        if self.accelerate and not self.initial_flag and self._plan is not None and self._plan[0] == (B, N, H, W) \
                and self._plan[1].workspace is not None:
            return self._plan[1]
        plan = self._prepare(input)
        self.initial_flag = not self.accelerate
        return plan

    def view_transform(self, input, depth, tran_feat, height, feat_nhwc=None, return_act=False):
        """depth (B*N, D, fH, fW), tran_feat (B*N, C, fH, fW), height (B*N, H, fH, fW) ->
        (bev, depth, height, low, mid, high) like LH:407-459; the four BEV tensors have the
        reference's logical shape (B, dz*C, Dy, Dx) [collapse_z] in `out_layout` memory.
        return_act=True (inference, bf16 speed mode, collapse_z=True): the four BEV tensors come back as bf16 NHWC
        activations (dhd_b200.dense.Act) written by the pool kernel itself."""
        B, N, _, H, W = input[0].shape
        if not depth.is_cuda:
            raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
        plan = self._bins(input)
        if feat_nhwc is None:
            feat_nhwc = tran_feat.view(B, N, self.out_channels, H, W).permute(0, 1, 3, 4, 2).contiguous()
        pixmask = height_to_mask(height, self.height_range, self.mask_range)
        if return_act:
            from dhd_b200 import dense as D
            if not self.collapse_z or D.PRECISIONS[self.precision][0] != 1:
                raise NotImplementedError('return_act=True: bf16 speed mode with collapse_z=True')
            outs = plan.alloc_outputs('nhwc_bf16', depth.device)
            plan.raw_forward(depth, feat_nhwc.view(B, N, H, W, self.out_channels), pixmask, outs, 'nhwc_bf16')
            outs = [D.Act(o, o.shape[-1], 1) for o in outs]
        elif self.collapse_z:
            outs = plan(depth, feat_nhwc.view(B, N, H, W, self.out_channels), pixmask, layout=self.out_layout)
            if self.out_layout == 'nhwc':
                outs = [o.permute(0, 3, 1, 2) for o in outs]        # logical NCHW, channels_last memory
        else:
            outs = plan(depth, feat_nhwc.view(B, N, H, W, self.out_channels), pixmask, layout='ncdhw')
        # the reference leaves the last pass's grid behind (LH:455; get_height_loss depends on it)
        self.grid_config = self.mask_3_grid
        self.create_grid_infos(**self.grid_config)
        return outs[0], depth, height, outs[1], outs[2], outs[3]

    def forward(self, input, stereo_metas=None, return_act=False):
        """input = [x (B,N,C,fH,fW), sensor2egos, ego2globals, intrins, post_rots, post_trans, bda,
        mlp_input] -> (bev, depth, height, low, mid, high), LH:461-490 (return_act: see view_transform)."""
        from dhd_b200 import dense as D
        from dhd_b200.modules import DepthHeadEngine
        x = input[0]
        mlp_input = input[7]
        B, N, C, H, W = x.shape
        if not x.is_cuda:
            raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
        from dhd_b200 import autograd as A
        if A.wants_grad(self, x):
            if return_act:
                raise NotImplementedError('return_act=True is the inference hand-off')
            # differentiable form (training): forward with saved activations, backward through the pool backward,
            # depth_net and -- via the height distribution's gradient (get_height_loss) -- HeightNet
            if stereo_metas is not None or self.height_net.stereo:
                raise NotImplementedError('MGHS: HeightNet is never called with a cost volume (lss_heightmap.py:787)')
            plan = self._bins(input)
            pix = lambda h: height_to_mask(h, self.height_range, self.mask_range)
            outs = A.mghs_forward(self, input, plan, pix)
            self.grid_config = self.mask_3_grid            # LH:455 quirk, as in view_transform
            self.create_grid_infos(**self.grid_config)
            return outs
        with torch.no_grad():
            xa = D.pack_input(x.reshape(B * N, C, H, W), D.PRECISIONS[self.precision][0])
            dev = x.device
            engine = self.cached_engine(dev, lambda: DepthHeadEngine(self.depth_net, self.D, self.precision, dev),
                                        slot='_depth_engine')
            depth, feat = engine(xa)                                   # softmax-ed depth, NHWC context
            height = self.height_net(xa, mlp_input, stereo_metas, softmax=True)
            return self.view_transform(input, depth, None, height, feat_nhwc=feat, return_act=return_act)

    # ------------------------------------------------------------------ mlp input (LH:493-526)
    def get_mlp_input(self, sensor2ego, ego2global, intrin, post_rot, post_tran, bda):
        B, N = sensor2ego.shape[:2]
        bda = bda.view(B, 1, 3, 3).repeat(1, N, 1, 1)
        cols = [intrin[:, :, 0, 0], intrin[:, :, 1, 1], intrin[:, :, 0, 2], intrin[:, :, 1, 2],
                post_rot[:, :, 0, 0], post_rot[:, :, 0, 1], post_tran[:, :, 0],
                post_rot[:, :, 1, 0], post_rot[:, :, 1, 1], post_tran[:, :, 1],
                bda[:, :, 0, 0], bda[:, :, 0, 1], bda[:, :, 1, 0], bda[:, :, 1, 1], bda[:, :, 2, 2]]
        return torch.cat([torch.stack(cols, dim=-1), sensor2ego[:, :, :3, :].reshape(B, N, -1)], dim=-1)

    # ------------------------------------------------------------------ height loss (LH:566-701)
    def _min_pool_sparse(self, maps):
        """downsample x downsample min-pool of a sparse map, zeros ignored (LH:633-645, 672-683)."""
        B, N, H, W = maps.shape
        ds = self.downsample
        t = torch.where(maps == 0.0, torch.full_like(maps, 1e5), maps)
        t = t.view(B * N, H // ds, ds, W // ds, ds).permute(0, 1, 3, 2, 4).reshape(-1, ds * ds)
        t = t.min(dim=-1).values
        return t

    def downsample_sparse_map(self, height_maps, downsample_factor=16):
        """LH:566-594: (B, N, H, W) sparse map -> (B, N, H/f, W/f) minimum of the non-zero values per f x f block,
        0 where a block is empty."""
        B, N, H, W = height_maps.shape
        f = downsample_factor
        assert H % f == 0 and W % f == 0, 'the map must be a whole number of blocks'
        t = torch.where(height_maps == 0.0, torch.full_like(height_maps, 1e5), height_maps)
        t = t.view(B, N, H // f, f, W // f, f).permute(0, 1, 2, 4, 3, 5).reshape(B, N, -1, f * f).min(dim=-1).values
        return torch.where(t == 1e5, torch.zeros_like(t), t).view(B, N, H // f, W // f)

    def get_downsampled_gt_depth(self, gt_depths):
        t = self._min_pool_sparse(gt_depths)        # empty blocks keep the 1e5 sentinel -> bin 0 below
        dc = self.grid_config['depth']
        if not self.sid:
            t = (t - (dc[0] - dc[2])) / dc[2]
        else:
            t = torch.log(t) - torch.log(torch.tensor(dc[0]).float())
            t = t * (self.D - 1) / torch.log(torch.tensor(dc[1] - 1.).float() / dc[0]) + 1.
        t = torch.where((t < self.D + 1) & (t >= 0.0), t, torch.zeros_like(t))
        return torch.nn.functional.one_hot(t.long(), num_classes=self.D + 1).view(-1, self.D + 1)[:, 1:].float()

    def get_downsampled_gt_height(self, gt_heights):
        t = self._min_pool_sparse(gt_heights)
        t = (t - self.height_range[0]) / self.height_interval     # LH:692 (no +1 offset, as the reference)
        t = torch.where((t < self.H + 1) & (t >= 0.0), t, torch.zeros_like(t))
        return torch.nn.functional.one_hot(t.long(), num_classes=self.H + 1).view(-1, self.H + 1)[:, 1:].float()

    @force_fp32()
    def get_height_loss(self, gt_depth, gt_height, height):
        """BCE between the height distribution and the binned LiDAR height on foreground pixels
        (pixels with a valid depth bin), LH:595-622."""
        labels = self.get_downsampled_gt_height(gt_height)
        fg = self.get_downsampled_gt_depth(gt_depth).max(dim=1).values > 0.0
        preds = height.permute(0, 2, 3, 1).contiguous().view(-1, self.H)
        loss = torch.nn.functional.binary_cross_entropy(preds[fg].float(), labels[fg], reduction='none').sum()
        return self.loss_height_weight * loss / max(1.0, float(fg.sum()))


@NECKS.register_module(force=True)
class MGHS_Depth(MGHS):
    """MGHS with the camera-aware DepthNet (DHD-M / DHD-L), reference LH:704-897.  Returns
    (bev, bev_w_z, depth, height); with collapse_z=False bev is (B, C, 1, Dy, Dx) and bev_w_z the
    low / mid / high slabs stacked on z, (B, C, 16, Dy, Dx) -- written in place by the pool kernel
    (DHD_LAYOUT_NCDHW_CAT) instead of three tensors + torch.cat."""

    def __init__(self, loss_depth_weight=3.0, depthnet_cfg=dict(), **kwargs):
        super().__init__(**kwargs)
        self.loss_depth_weight = loss_depth_weight
        self.depth_net = DepthNet(in_channels=self.in_channels, mid_channels=self.in_channels,
                                  context_channels=self.out_channels, depth_channels=self.D,
                                  precision=self.precision, **depthnet_cfg)

    def forward(self, input, stereo_metas=None, return_act=False):
        """return_act=True (inference, bf16 speed mode, collapse_z=False): bev / bev_w_z come back as bf16 NHWC
        activations (dhd_b200.dense.Act) with the z planes collapsed into channels (channel = z*C + c, what the
        detectors' `torch.cat(x.unbind(dim=2), 1)` produces) -- the pool kernel writes them in that form directly."""
        from dhd_b200 import dense as D
        x, mlp_input = input[0], input[7]
        B, N, C, H, W = x.shape
        from dhd_b200 import autograd as A
        if x.is_cuda and A.wants_grad(self, x):
            if return_act:
                raise NotImplementedError('return_act=True is the inference hand-off')
            # differentiable form (training): DepthNetTrainer + HeightNetTrainer + pool backward (dhd_b200.autograd)
            plan = self._bins(input)
            pix = lambda h: height_to_mask(h, self.height_range, self.mask_range)
            outs = A.mghs_depth_forward(self, input, stereo_metas, plan, pix)
            self.grid_config = dict(_BEV_PASS_GRID)        # "reset grid_config!", LH:848-854
            self.create_grid_infos(**self.grid_config)
            return outs
        if not x.is_cuda:
            raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
        with torch.no_grad():
            xa = D.pack_input(x.reshape(B * N, C, H, W), D.PRECISIONS[self.precision][0])
            depth, feat = self.depth_net.forward_split(xa, mlp_input, softmax=True, stereo_metas=stereo_metas)
            height = self.height_net(xa, mlp_input, None, softmax=True)
            return self.view_transform(input, depth, None, height, feat_nhwc=feat, return_act=return_act)

    def view_transform(self, input, depth, tran_feat, height, feat_nhwc=None, return_act=False):
        B, N, _, H, W = input[0].shape
        if not depth.is_cuda:
            raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
        plan = self._bins(input)
        if feat_nhwc is None:
            feat_nhwc = tran_feat.view(B, N, self.out_channels, H, W).permute(0, 1, 3, 4, 2).contiguous()
        feat_nhwc = feat_nhwc.view(B, N, H, W, self.out_channels)
        pixmask = height_to_mask(height, self.height_range, self.mask_range)
        if return_act:
            from dhd_b200 import dense as D
            if self.collapse_z or D.PRECISIONS[self.precision][0] != 1:
                raise NotImplementedError('return_act=True: bf16 speed mode with collapse_z=False')
            outs = plan.alloc_outputs('nhwc_bf16', depth.device)           # four (B, Dy, Dx, dz*C) bf16 tensors
            plan.raw_forward(depth, feat_nhwc, pixmask, outs, 'nhwc_bf16')
            w_z = torch.cat(outs[1:], dim=-1)                              # low | mid | high slabs stacked on z
            bev, bev_w_z = D.Act(outs[0], outs[0].shape[-1], 1), D.Act(w_z, w_z.shape[-1], 1)
        elif self.collapse_z:       # not used by any config; same concatenation on dim 2 as the reference
            outs = plan(depth, feat_nhwc, pixmask, layout='nchw')
            bev, bev_w_z = outs[0], torch.cat(outs[1:], dim=2)
        else:
            bev, bev_w_z = plan(depth, feat_nhwc, pixmask, layout='ncdhw_cat')
        self.grid_config = dict(_BEV_PASS_GRID)        # "reset grid_config!", LH:848-854
        self.create_grid_infos(**self.grid_config)
        return bev, bev_w_z, depth, height

    @force_fp32()
    def get_depth_and_height_loss(self, gt_depth, gt_height, depth, height):
        """LH:859-897: BCE of the depth and height distributions on foreground pixels."""
        hl = self.get_downsampled_gt_height(gt_height)
        dl = self.get_downsampled_gt_depth(gt_depth)
        fg = dl.max(dim=1).values > 0.0
        hp = height.permute(0, 2, 3, 1).contiguous().view(-1, self.H)[fg]
        dp = depth.permute(0, 2, 3, 1).contiguous().view(-1, self.D)[fg]
        bce = torch.nn.functional.binary_cross_entropy
        norm = max(1.0, float(fg.sum()))
        return (self.loss_depth_weight * bce(dp.float(), dl[fg], reduction='none').sum() / norm,
                self.loss_height_weight * bce(hp.float(), hl[fg], reduction='none').sum() / norm)


@NECKS.register_module(force=True)
class MGHS_Stereo(MGHS_Depth):
    """LH:900-907: MGHS_Depth + the 1/4-resolution frustum template the stereo cost volume uses."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.cv_frustum = self.create_frustum(kwargs['grid_config']['depth'], kwargs['input_size'], downsample=4)
