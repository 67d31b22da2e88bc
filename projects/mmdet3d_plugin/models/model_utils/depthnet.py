"""HeightNet and its building blocks: parameter containers with the reference's attribute names
(checkpoint-compatible state_dict) whose forward pass runs on the
B200 engines in dhd_b200.modules -- tcgen05 implicit-GEMM convolutions with fused epilogues.

Reference: projects/mmdet3d_plugin/models/model_utils/depthnet.py
  _ASPPModule / ASPP 10-116, Mlp 119-147, SELayer 150-169, HeightNet 418-487 + forward 605-652.
Initialisation follows the reference (kaiming_normal for ASPP convs, BN weight 1 / bias 0).
"""
import torch
import torch.nn as nn

from dhd_b200.compat import BasicBlock, build_conv_layer


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


class HeightNet(nn.Module):
    """Per-pixel height distribution head (reference depthnet.py:418-652)."""

    def __init__(self, in_channels, mid_channels, depth_channels, use_dcn=True, use_aspp=True,
                 with_cp=False, stereo=False, bias=0.0, aspp_mid_channels=-1, precision='fp32'):
        super().__init__()
        if stereo:
            raise NotImplementedError('HeightNet(stereo=True) cost-volume branch: SURVEY 8(f) rank 3')
        self.reduce_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, stride=1, padding=1),
            nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True))
        self.bn = nn.BatchNorm1d(27)
        self.depth_mlp = Mlp(27, mid_channels, mid_channels)
        self.depth_se = SELayer(mid_channels)
        layers = [BasicBlock(mid_channels, mid_channels) for _ in range(3)]
        if use_aspp:
            layers.append(ASPP(mid_channels, mid_channels if aspp_mid_channels < 0 else aspp_mid_channels))
        if use_dcn:
            layers.append(build_conv_layer(cfg=dict(type='DCN', in_channels=mid_channels,
                                                    out_channels=mid_channels, kernel_size=3, padding=1,
                                                    groups=4, im2col_step=128)))
        layers.append(nn.Conv2d(mid_channels, depth_channels, kernel_size=1, stride=1, padding=0))
        self.depth_conv = nn.Sequential(*layers)
        self.with_cp = with_cp
        self.depth_channels = depth_channels
        self.precision = precision
        self._engine = None

    def engine(self, device):
        from dhd_b200.modules import HeightNetEngine
        if self._engine is None or self._engine.device != device:
            self._engine = HeightNetEngine(self, self.precision, device)
        return self._engine

    def invalidate(self):
        """Call after changing parameters (load_state_dict does it automatically)."""
        self._engine = None

    def _load_from_state_dict(self, *a, **k):
        self._engine = None
        return super()._load_from_state_dict(*a, **k)

    def train(self, mode=True):
        self._engine = None
        return super().train(mode)

    def forward(self, x, mlp_input, stereo_metas=None, softmax=False):
        """x: (B*N, C, fH, fW) fp32 CUDA tensor or a dhd_b200.dense.Act; mlp_input (B, N, 27).
        Returns the height logits (B*N, H, fH, fW) like the reference (softmax=True fuses the
        channel softmax MGHS.forward applies next into the last layer's epilogue)."""
        from dhd_b200 import dense as D
        if self.training:
            raise NotImplementedError('dhd_b200 HeightNet: inference (eval-mode BatchNorm) only in this build')
        if stereo_metas is not None:
            raise NotImplementedError('stereo_metas: SURVEY 8(f) rank 3')
        if not isinstance(x, D.Act):
            if not x.is_cuda:
                raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
            x = D.pack_input(x, D.PRECISIONS[self.precision][0])
        with torch.no_grad():
            return self.engine(x.data.device)(x, mlp_input, softmax=softmax)
