"""TEST INFRASTRUCTURE -- torch fp32 restatements of the third-party layers the
reference's dense hot-path modules pull from un-vendored packages:

  mmcv-full 1.5.3  DCN == DeformConv2dPack (depthnet.py:225-236, 466-477)
  mmdet 2.25.1     BasicBlock               (depthnet.py:4, 217-220, 458-461)
  mmcv-full 1.5.3  ConvModule               (occ_head.py:52-60)

PARITY UNPINNED for these three: the reference ships no test vector at these
boundaries and the packages are not installable here; the restatements follow the
published layer definitions (DeformConv2dPack: zero-initialised 3x3 offset conv
producing deform_groups*2*kh*kw channels in (dy, dx) interleaved order, then
deformable conv without bias).
"""
import torch
import torch.nn as nn


class DeformConv2dPack(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, deform_groups=1, bias=False, **kw):
        super().__init__()
        k = kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deform_groups = groups, deform_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, k, k))
        nn.init.kaiming_uniform_(self.weight, nonlinearity='relu')
        self.conv_offset = nn.Conv2d(in_channels, deform_groups * 2 * k * k, k, stride=stride,
                                     padding=padding, dilation=dilation, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        from torchvision.ops import deform_conv2d
        off = self.conv_offset(x)
        return deform_conv2d(x, off, self.weight, None, stride=self.stride, padding=self.padding,
                             dilation=self.dilation)
