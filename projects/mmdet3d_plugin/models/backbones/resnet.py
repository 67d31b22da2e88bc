"""CustomResNet BEV encoder backbone (reference: projects/mmdet3d_plugin/models/backbones/resnet.py:10-80,
block_type='Basic'): stages of mmdet BasicBlocks, the first block of a stage stride 2 with a 3x3 stride-2
conv as `downsample`.  Same parameter names (`layers.<stage>.<block>.conv1 / bn1 / conv2 / bn2 / downsample`),
forward on dhd_b200.encoders.CustomResNetEngine."""
import torch
import torch.nn as nn

from dhd_b200.compat import BACKBONES, BasicBlock, EngineOwner


@BACKBONES.register_module(force=True)
class CustomResNet(EngineOwner, nn.Module):
    def __init__(self, numC_input, num_layer=[2, 2, 2], num_channels=None, stride=[2, 2, 2],
                 backbone_output_ids=None, norm_cfg=dict(type='BN'), with_cp=False, block_type='Basic',
                 precision='fp32'):
        super().__init__()
        if block_type != 'Basic':
            raise NotImplementedError("CustomResNet(block_type='BottleNeck') is not used by the DHD configs")
        assert len(num_layer) == len(stride)
        num_channels = [numC_input * 2 ** (i + 1) for i in range(len(num_layer))] if num_channels is None \
            else num_channels
        self.backbone_output_ids = range(len(num_layer)) if backbone_output_ids is None else backbone_output_ids
        layers, cur = [], numC_input
        for i in range(len(num_layer)):
            blocks = [BasicBlock(cur, num_channels[i], stride=stride[i],
                                 downsample=nn.Conv2d(cur, num_channels[i], 3, stride[i], 1))]
            cur = num_channels[i]
            blocks += [BasicBlock(cur, cur, stride=1, downsample=None) for _ in range(num_layer[i] - 1)]
            layers.append(nn.Sequential(*blocks))
        self.layers = nn.Sequential(*layers)
        self.with_cp, self.precision = with_cp, precision
        self._engine = None

    def forward(self, x, return_act=False):
        """x: (B, C, Dy, Dx) -> list of (B, C_i, Dy/2^(i+1), Dx/2^(i+1)) (Acts with return_act=True)."""
        from dhd_b200 import dense as D
        from dhd_b200.encoders import CustomResNetEngine
        from dhd_b200.modules import unpack
        from dhd_b200 import autograd as A
        if not isinstance(x, D.Act) and x.is_cuda and A.wants_grad(self, x):
            return A.resnet_forward(self, x)         # differentiable form (dhd_b200.autograd)
        with torch.no_grad():
            if not isinstance(x, D.Act):
                if not x.is_cuda:
                    raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
                x = D.pack_any(x, D.PRECISIONS[self.precision][0])
            dev = x.data.device
            feats = self.cached_engine(dev, lambda: CustomResNetEngine(self, self.precision, dev))(x)
            return feats if return_act else [unpack(f) for f in feats]
