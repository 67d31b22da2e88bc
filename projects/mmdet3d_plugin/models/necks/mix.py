"""SFA dual-feature fusion neck (reference: projects/mmdet3d_plugin/models/necks/mix.py:8-90):
same parameter names (mysk_7.fc / mysk_7.spacial_leanring / mix_residual / mix_shortcut),
forward on the B200 engine (dhd_b200.modules.SFAEngine): squeeze + fc gate on CUDA cores, the
two gated blends as streaming kernels, the five convolutions on tcgen05 with BN / ReLU /
sigmoid / residual fused into their epilogues."""
import torch
import torch.nn as nn

from dhd_b200.compat import NECKS, BaseModule, EngineOwner


class channel_spatial_stage(nn.Module):
    def __init__(self, features):
        super().__init__()
        reduction = 16
        self.channels = features // 2
        self.fc = nn.Sequential(nn.Linear(features, features // reduction), nn.ReLU(inplace=False),
                                nn.Linear(features // reduction, self.channels), nn.Sigmoid())
        self.spacial_leanring = nn.Sequential(
            nn.Conv2d(self.channels, self.channels, kernel_size=1, stride=1, padding=0),
            nn.BatchNorm2d(self.channels), nn.ReLU(inplace=True),
            nn.Conv2d(self.channels, self.channels, kernel_size=1, stride=1, padding=0),
            nn.BatchNorm2d(self.channels))
        self.sigmoid = nn.Sigmoid()


@NECKS.register_module(force=True)
class SFA(EngineOwner, BaseModule):
    def __init__(self, in_channels, out_channels, stride=1, precision='fp32'):
        super().__init__()
        if stride != 1:
            raise NotImplementedError('SFA stride != 1 is not used by any DHD config')
        self.mysk_7 = channel_spatial_stage(features=in_channels)
        self.mix_channels, self.out_channels = in_channels, out_channels
        self.mix_residual = nn.Sequential(
            nn.Conv2d(in_channels // 2, out_channels, kernel_size=3, stride=stride, padding=1, bias=False),
            nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(out_channels))
        self.mix_shortcut = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, stride=stride, kernel_size=1, bias=False),
            nn.BatchNorm2d(out_channels))
        self.relu = nn.ReLU(inplace=True)
        self.precision = precision
        self._engine = None

    def forward(self, inputs, return_act=False):
        """inputs: (B, 2C, Dy, Dx) fp32 CUDA tensor (any memory format) or a dhd_b200.dense.Act.
        Returns (B, C_out, Dy, Dx) fp32 (logical NCHW, channels_last memory), or the Act when
        return_act=True (what predictor.forward consumes without a layout round trip).
        Under autograd (a parameter or the input requires grad) the call runs the training engine and is
        differentiable (dhd_b200.autograd): BatchNorm on batch statistics in train(), frozen in eval()."""
        from dhd_b200 import autograd as A
        from dhd_b200 import dense as D
        from dhd_b200.modules import SFAEngine
        if not isinstance(inputs, D.Act):
            if not inputs.is_cuda:
                raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
            if A.wants_grad(self, inputs):
                return A.sfa_forward(self, inputs)
        with torch.no_grad():
            if not isinstance(inputs, D.Act):
                inputs = D.pack_any(inputs, D.PRECISIONS[self.precision][0], want_mean=True)
            dev = inputs.data.device
            engine = self.cached_engine(dev, lambda: SFAEngine(self, self.precision, dev))
            if return_act:
                return engine(inputs)
            N, H, W = inputs.N, inputs.H, inputs.W
            out = torch.empty(N, H, W, self.out_channels, device=dev)
            engine(inputs, out_f32=(out, D.nhwc_strides(self.out_channels, H, W)))
            return out.permute(0, 3, 1, 2)
