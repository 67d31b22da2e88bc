"""UNet voxel encoder (reference: projects/mmdet3d_plugin/models/backbones/unet.py:6-141): parameter containers
with the reference's attribute names (inc / down1-4 / up1-4 / outc, `double_conv`, `maxpool_conv`, `up`, `conv`),
forward on the B200 engine (dhd_b200.encoders.UNetEngine)."""
import torch
import torch.nn as nn

from dhd_b200.compat import BACKBONES, EngineOwner


class DoubleConv(nn.Module):
    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid_channels = mid_channels or out_channels
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(mid_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))


class Down(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))


class Up(nn.Module):
    def __init__(self, in_channels, out_channels, bilinear=False):
        super().__init__()
        if bilinear:
            raise NotImplementedError('UNet(bilinear=True) is not used by the DHD configs')
        self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(in_channels, out_channels)


class OutConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)


@BACKBONES.register_module(force=True)
class UNet(EngineOwner, nn.Module):
    def __init__(self, n_channels, n_classes, bilinear=False, precision='fp32'):
        super().__init__()
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, bilinear
        self.inc = DoubleConv(n_channels, 64)
        self.down1, self.down2, self.down3 = Down(64, 128), Down(128, 256), Down(256, 512)
        self.down4 = Down(512, 1024)
        self.up1, self.up2 = Up(1024, 512, bilinear), Up(512, 256, bilinear)
        self.up3, self.up4 = Up(256, 128, bilinear), Up(128, 64, bilinear)
        self.outc = OutConv(64, n_classes)
        self.precision = precision
        self._engine = None

    def forward(self, x, return_act=False, out=None):
        """x: (B, n_channels, H, W) fp32 CUDA tensor (any memory format) or a dhd_b200.dense.Act ->
        (B, n_classes, H, W) fp32 (channels_last memory), or the Act with return_act=True."""
        from dhd_b200 import dense as D
        from dhd_b200.encoders import UNetEngine
        from dhd_b200.modules import unpack
        from dhd_b200 import autograd as A
        if not isinstance(x, D.Act) and x.is_cuda and A.wants_grad(self, x):
            return A.unet_forward(self, x)           # differentiable form (dhd_b200.autograd)
        with torch.no_grad():
            if not isinstance(x, D.Act):
                if not x.is_cuda:
                    raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
                x = D.pack_any(x, D.PRECISIONS[self.precision][0])
            dev = x.data.device
            y = self.cached_engine(dev, lambda: UNetEngine(self, self.precision, dev))(x, out=out)
            return y if return_act else unpack(y.slice(0, self.n_classes))
