"""FPN_LSS BEV encoder neck (reference: projects/mmdet3d_plugin/models/necks/lss_fpn.py:11-74): bilinear x4
up-sampling of the coarsest map, concat with the finest, two conv3x3-BN-ReLU, bilinear x2, conv3x3-BN-ReLU,
conv1x1.  Same parameter names (`conv.{0,1,3,4}`, `up2.{1,2,4}`), forward on dhd_b200.encoders.FPNLSSEngine."""
import torch
import torch.nn as nn

from dhd_b200.compat import NECKS, EngineOwner


@NECKS.register_module(force=True)
class FPN_LSS(EngineOwner, nn.Module):
    def __init__(self, in_channels, out_channels, scale_factor=4, input_feature_index=(0, 2), norm_cfg=dict(type='BN'),
                 extra_upsample=2, lateral=None, use_input_conv=False, precision='fp32'):
        super().__init__()
        if lateral is not None:
            raise NotImplementedError('FPN_LSS(lateral=...) is not used by the DHD configs')
        self.input_feature_index = input_feature_index
        self.extra_upsample = extra_upsample is not None
        self.out_channels = out_channels
        self.lateral = False
        self.up = nn.Upsample(scale_factor=scale_factor, mode='bilinear', align_corners=True)
        f = 2 if self.extra_upsample else 1
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels * f, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_channels * f),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels * f, out_channels * f, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(out_channels * f), nn.ReLU(inplace=True))
        if self.extra_upsample:
            self.up2 = nn.Sequential(
                nn.Upsample(scale_factor=extra_upsample, mode='bilinear', align_corners=True),
                nn.Conv2d(out_channels * f, out_channels, kernel_size=3, padding=1, bias=False),
                nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True),
                nn.Conv2d(out_channels, out_channels, kernel_size=1, padding=0))
        self.precision = precision
        self._engine = None

    def forward(self, feats, return_act=False, out=None):
        """feats: list of (B, C_i, H_i, W_i) tensors or Acts -> (B, out_channels, 2H, 2W)."""
        from dhd_b200 import dense as D
        from dhd_b200.encoders import FPNLSSEngine
        from dhd_b200.modules import unpack
        from dhd_b200 import autograd as A
        if all(not isinstance(f, D.Act) and f.is_cuda for f in feats) and A.wants_grad(self, *feats):
            return A.fpn_forward(self, list(feats))  # differentiable form (dhd_b200.autograd)
        with torch.no_grad():
            parts = D.PRECISIONS[self.precision][0]
            acts = []
            for f in feats:
                if not isinstance(f, D.Act):
                    if not f.is_cuda:
                        raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
                    f = D.pack_any(f, parts)
                acts.append(f)
            dev = acts[0].data.device
            y = self.cached_engine(dev, lambda: FPNLSSEngine(self, self.precision, dev))(acts, out=out)
            return y if return_act else unpack(y)
