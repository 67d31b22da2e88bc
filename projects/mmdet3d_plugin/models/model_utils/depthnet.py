"""HeightNet and its building blocks: parameter containers with the reference's attribute names
(checkpoint-compatible state_dict) whose forward pass runs on the
B200 engines in dhd_b200.modules -- tcgen05 implicit-GEMM convolutions with fused epilogues.

Reference: projects/mmdet3d_plugin/models/model_utils/depthnet.py
  _ASPPModule / ASPP 10-116, Mlp 119-147, SELayer 150-169, HeightNet 418-487 + forward 605-652.
Initialisation follows the reference (kaiming_normal for ASPP convs, BN weight 1 / bias 0).
"""
import torch
import torch.nn as nn

from dhd_b200.compat import BasicBlock, EngineOwner, build_conv_layer


class _ASPPModule(nn.Module):
    def __init__(self, inplanes, planes, kernel_size, padding, dilation):
        super().__init__()
        self.atrous_conv = nn.Conv2d(inplanes, planes, kernel_size, stride=1, padding=padding,
                                     dilation=dilation, bias=False)
        self.bn = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU()
        nn.init.kaiming_normal_(self.atrous_conv.weight)


class ASPP(nn.Module):
    """1x1 + three dilated 3x3 branches (6 / 12 / 18) + global-pool branch -> 1x1 -> BN -> ReLU ->
    Dropout(0.5) (identity in eval)."""

    def __init__(self, inplanes, mid_channels=256):
        super().__init__()
        self.aspp1 = _ASPPModule(inplanes, mid_channels, 1, 0, 1)
        self.aspp2 = _ASPPModule(inplanes, mid_channels, 3, 6, 6)
        self.aspp3 = _ASPPModule(inplanes, mid_channels, 3, 12, 12)
        self.aspp4 = _ASPPModule(inplanes, mid_channels, 3, 18, 18)
        self.global_avg_pool = nn.Sequential(
            nn.AdaptiveAvgPool2d((1, 1)), nn.Conv2d(inplanes, mid_channels, 1, stride=1, bias=False),
            nn.BatchNorm2d(mid_channels), nn.ReLU())
        self.conv1 = nn.Conv2d(mid_channels * 5, inplanes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.ReLU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)


class SELayer(nn.Module):
    def __init__(self, channels, act_layer=nn.ReLU, gate_layer=nn.Sigmoid):
        super().__init__()
        self.conv_reduce = nn.Conv2d(channels, channels, 1, bias=True)
        self.act1 = act_layer()
        self.conv_expand = nn.Conv2d(channels, channels, 1, bias=True)
        self.gate = gate_layer()


def _trunk_layers(mid_channels, depth_channels, use_dcn, use_aspp, aspp_mid_channels, stereo):
    """depth_conv of DepthNet / HeightNet: 3 BasicBlocks [+ ASPP] [+ DCN] + 1x1 (depthnet.py:203-243, 448-484)."""
    first_in, downsample = mid_channels, None
    if stereo:
        first_in = mid_channels + depth_channels
        downsample = nn.Conv2d(first_in, mid_channels, 1, 1, 0)
    layers = [BasicBlock(first_in, mid_channels, downsample=downsample),
              BasicBlock(mid_channels, mid_channels), BasicBlock(mid_channels, mid_channels)]
    if use_aspp:
        layers.append(ASPP(mid_channels, mid_channels if aspp_mid_channels < 0 else aspp_mid_channels))
    if use_dcn:
        layers.append(build_conv_layer(cfg=dict(type='DCN', in_channels=mid_channels, out_channels=mid_channels,
                                                kernel_size=3, padding=1, groups=4, im2col_step=128)))
    layers.append(nn.Conv2d(mid_channels, depth_channels, kernel_size=1, stride=1, padding=0))
    return nn.Sequential(*layers)


def _cost_volumn_net(depth_channels):
    layers = []
    for _ in range(2):
        layers += [nn.Conv2d(depth_channels, depth_channels, kernel_size=3, stride=2, padding=1),
                   nn.BatchNorm2d(depth_channels)]
    return nn.Sequential(*layers)


class _PlaneSweep:
    """The two stereo helpers DepthNet and HeightNet share in the reference (depthnet.py:245-361, 489-603)."""

    def gen_grid(self, metas, B, N, D, H, W, hi, wi):
        """Reference depthnet.py:245-308: the (B*N, D*H, W, 2) normalised sampling coordinates of the previous frame
        for every point of the current frame's 1/4-resolution frustum (points behind the previous camera -> -2).
        Plain torch, for callers of the reference API; `calculate_cost_volumn` computes the same coordinates inside
        its kernel (bit-equal, tests/test_stereo_gpu.py) and never materialises this tensor."""
        frustum = metas['frustum']
        view = lambda t, *tail: t.view(B, N, 1, 1, 1, *tail)
        pts = frustum - view(metas['post_trans'], 3)
        pts = view(torch.inverse(metas['post_rots']), 3, 3).matmul(pts.unsqueeze(-1))
        pts = torch.cat((pts[..., :2, :] * pts[..., 2:3, :], pts[..., 2:3, :]), 5)          # (u d, v d, d)
        k2s = metas['k2s_sensor']
        combine = k2s[:, :, :3, :3].contiguous().matmul(torch.inverse(metas['intrins']))
        pts = view(combine, 3, 3).matmul(pts)
        pts = pts + view(k2s[:, :, :3, 3].contiguous(), 3, 1)                               # previous camera frame
        behind = pts[..., 2, 0] < 1e-3
        pts = view(metas['intrins'], 3, 3).matmul(pts)
        pts = pts[..., :2, :] / pts[..., 2:3, :]
        pts = view(metas['post_rots'][..., :2, :2], 2, 2).matmul(pts).squeeze(-1)
        pts = pts + view(metas['post_trans'][..., :2], 2)
        px = pts[..., 0] / (wi - 1.0) * 2.0 - 1.0
        py = pts[..., 1] / (hi - 1.0) * 2.0 - 1.0
        px[behind] = -2
        py[behind] = -2
        return torch.stack([px, py], dim=-1).view(B * N, D * H, W, 2)

    def calculate_cost_volumn(self, metas, out_act=None):
        """Reference depthnet.py:310-361 (+ gen_grid 245-308) as ONE fused kernel (dhd_b200/csrc/stereo.cu):
        (B*N, D, fH_stereo, fW_stereo) matching probabilities from stereo_metas (same dict as the reference).
        out_act: optional split-bf16 activation to fill instead (the form cost_volumn_net reads)."""
        from dhd_b200 import stereo as S
        prev, curr = metas['cv_feat_list']
        if not curr.is_cuda:
            raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
        frustum = metas['frustum']
        D, H, W, _ = frustum.shape
        lowp = self.precision == 'bf16'
        with torch.no_grad():
            cam = S.camera_table(metas['k2s_sensor'], metas['intrins'], metas['post_rots'], metas['post_trans'])
            p = S.to_nhwc(prev.reshape(-1, prev.shape[-3], H, W).float(), bf16=lowp)
            c = S.to_nhwc(curr.reshape(-1, curr.shape[-3], H, W).float(), bf16=lowp)
            # gen_grid's hi, wi = 4 * (stereo map size), depthnet.py:339-340
            out, _ = S.cost_volume(p, c, D, (H * 4, W * 4), bias=self.bias, frustum=frustum.to(curr.device), cam=cam,
                                   out_act=out_act)
        return out


class DepthNet(EngineOwner, _PlaneSweep, nn.Module):
    """Depth + context head of MGHS_Depth / MGHS_Stereo (reference depthnet.py:172-415).
    stereo=True adds cost_volumn_net and the first block's 1x1 downsample; the plane-sweep cost volume
    (gen_grid + calculate_cost_volumn, 245-361) is one fused CUDA kernel, see calculate_cost_volumn."""

    def __init__(self, in_channels, mid_channels, context_channels, depth_channels, use_dcn=True,
                 use_aspp=True, with_cp=False, stereo=False, bias=0.0, aspp_mid_channels=-1, precision='fp32'):
        super().__init__()
        self.reduce_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, stride=1, padding=1),
            nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True))
        self.context_conv = nn.Conv2d(mid_channels, context_channels, kernel_size=1, stride=1, padding=0)
        self.bn = nn.BatchNorm1d(27)
        self.depth_mlp = Mlp(27, mid_channels, mid_channels)
        self.depth_se = SELayer(mid_channels)
        self.context_mlp = Mlp(27, mid_channels, mid_channels)
        self.context_se = SELayer(mid_channels)
        if stereo:
            self.cost_volumn_net = _cost_volumn_net(depth_channels)
        self.bias = bias
        self.depth_conv = _trunk_layers(mid_channels, depth_channels, use_dcn, use_aspp, aspp_mid_channels, stereo)
        self.with_cp, self.depth_channels, self.stereo = with_cp, depth_channels, stereo
        self.precision = precision
        self._engine = None

    def forward_split(self, x, mlp_input, softmax=True, stereo_metas=None):
        """-> (depth (B*N, D, fH, fW) NCHW [softmax-ed], context (B*N, fH, fW, C) NHWC): the layouts the
        fused pool consumes, written directly by the layer epilogues."""
        from dhd_b200 import dense as D
        from dhd_b200.modules import DepthNetEngine
        if self.stereo != (stereo_metas is not None):
            # the reference fails the same way: the first BasicBlock's input width depends on the cost volume
            raise RuntimeError('DepthNet(stereo=%s) called %s stereo_metas' %
                               (self.stereo, 'with' if stereo_metas is not None else 'without'))
        if not isinstance(x, D.Act):
            if not x.is_cuda:
                raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
            x = D.pack_input(x, D.PRECISIONS[self.precision][0])
        with torch.no_grad():
            dev = x.data.device
            engine = self.cached_engine(dev, lambda: DepthNetEngine(self, self.precision, dev))
            cv = None
            if self.stereo:
                # depthnet.py:387-400: zeros when there is no previous frame, else the plane-sweep volume
                scale = float(stereo_metas['downsample']) / stereo_metas['cv_downsample']
                Hs, Ws = int(x.H * scale), int(x.W * scale)
                first = stereo_metas['cv_feat_list'][0] is None
                cv = engine.new_cost_volume(x.N, Hs, Ws, x.data.device, zero=first)
                if not first:
                    self.calculate_cost_volumn(stereo_metas, out_act=cv)
            return engine(x, mlp_input, softmax=softmax, cost_volume=cv)

    def forward(self, x, mlp_input, stereo_metas=None):
        """Reference signature: (B*N, D + C_context, fH, fW) logits, depthnet.py:362-415."""
        depth, ctx = self.forward_split(x, mlp_input, softmax=False, stereo_metas=stereo_metas)
        return torch.cat([depth, ctx.permute(0, 3, 1, 2)], dim=1)


class HeightNet(EngineOwner, _PlaneSweep, nn.Module):
    """Per-pixel height distribution head (reference depthnet.py:418-652)."""

    def __init__(self, in_channels, mid_channels, depth_channels, use_dcn=True, use_aspp=True,
                 with_cp=False, stereo=False, bias=0.0, aspp_mid_channels=-1, precision='fp32'):
        super().__init__()
        self.stereo = stereo
        self.reduce_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, stride=1, padding=1),
            nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True))
        self.bn = nn.BatchNorm1d(27)
        self.depth_mlp = Mlp(27, mid_channels, mid_channels)
        self.depth_se = SELayer(mid_channels)
        if stereo:
            self.cost_volumn_net = _cost_volumn_net(depth_channels)
        self.bias = bias
        self.depth_conv = _trunk_layers(mid_channels, depth_channels, use_dcn, use_aspp, aspp_mid_channels, stereo)
        self.with_cp = with_cp
        self.depth_channels = depth_channels
        self.precision = precision
        self._engine = None

    def engine(self, device):
        from dhd_b200.modules import HeightNetEngine
        return self.cached_engine(device, lambda: HeightNetEngine(self, self.precision, device))

    def forward(self, x, mlp_input, stereo_metas=None, softmax=False):
        """x: (B*N, C, fH, fW) fp32 CUDA tensor or a dhd_b200.dense.Act; mlp_input (B, N, 27).
        Returns the height logits (B*N, H, fH, fW) like the reference (softmax=True fuses the
        channel softmax MGHS.forward applies next into the last layer's epilogue)."""
        from dhd_b200 import dense as D
        if stereo_metas is not None or self.stereo:
            raise NotImplementedError('HeightNet with a cost volume: no DHD module calls it that way '
                                      '(MGHS_Depth passes stereo_metas=None, lss_heightmap.py:787)')
        if not isinstance(x, D.Act):
            if not x.is_cuda:
                raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
            x = D.pack_input(x, D.PRECISIONS[self.precision][0])
        with torch.no_grad():
            return self.engine(x.data.device)(x, mlp_input, softmax=softmax)
