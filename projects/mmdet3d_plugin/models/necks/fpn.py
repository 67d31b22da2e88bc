"""CustomFPN image neck (reference: projects/mmdet3d_plugin/models/necks/fpn.py:11-203; DHD-S.py:56-62
`dict(type='CustomFPN', in_channels=[1024, 2048], out_channels=256, num_outs=1, start_level=0, out_ids=[0])`): 1x1
lateral convs, top-down nearest up-sampling + add, 3x3 output conv on the levels in `out_ids`.  Same constructor kwargs and
parameter names (`lateral_convs.{i}.conv`, `fpn_convs.{j}.conv`); forward on dhd_b200.backbone.CustomFPNEngine, under
autograd on dhd_b200.train_backbone.CustomFPNTrainer (forward with saved activations + hand-written backward)."""
import torch
import torch.nn as nn

from dhd_b200.compat import NECKS, BaseModule, ConvModule, EngineOwner


@NECKS.register_module(force=True)
class CustomFPN(EngineOwner, BaseModule):
    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, out_ids=[], add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode='nearest'), init_cfg=None, precision='bf16'):
        super().__init__()
        assert isinstance(in_channels, list)
        if add_extra_convs or norm_cfg is not None or act_cfg is not None or upsample_cfg.get('mode', 'nearest') != 'nearest' \
                or 'scale_factor' in upsample_cfg:
            raise NotImplementedError('CustomFPN variant outside the DHD configs (no extra convs / norm / activation, '
                                      'nearest up-sampling to the finer level\'s size)')
        self.in_channels, self.out_channels, self.num_ins, self.num_outs = in_channels, out_channels, len(in_channels), num_outs
        self.backbone_end_level = self.num_ins if end_level == -1 else end_level
        self.start_level, self.end_level, self.out_ids = start_level, end_level, list(out_ids)
        self.add_extra_convs, self.precision = False, precision
        if num_outs > len(self.out_ids):
            raise NotImplementedError('CustomFPN extra max-pool levels are not used by the DHD configs')
        self.lateral_convs, self.fpn_convs = nn.ModuleList(), nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(in_channels[i], out_channels, 1, norm_cfg=None, act_cfg=None, inplace=False))
            if i in self.out_ids:
                self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1, norm_cfg=None, act_cfg=None,
                                                 inplace=False))
        self._engine = None

    def forward(self, inputs, return_act=False):
        """inputs: the backbone's feature maps (tensors or Acts) -> list with one (N, out_channels, H, W) map per out_id."""
        from dhd_b200 import dense as D
        from dhd_b200.backbone import CustomFPNEngine
        from dhd_b200.modules import unpack
        assert len(inputs) == len(self.in_channels)
        from dhd_b200 import autograd as A
        if not any(isinstance(t, D.Act) for t in inputs) and A.wants_grad(self, *inputs):
            outs = A.custom_fpn_forward(self, list(inputs))       # differentiable form (dhd_b200.train_backbone)
            return outs
        with torch.no_grad():
            acts = []
            for t in inputs:
                if not isinstance(t, D.Act):
                    if not t.is_cuda:
                        raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
                    t = D.pack_any(t, D.PRECISIONS[self.precision][0])
                acts.append(t)
            dev = acts[0].data.device
            outs = self.cached_engine(dev, lambda: CustomFPNEngine(self, self.precision, dev))(acts)
            return outs if return_act else [unpack(o) for o in outs]
