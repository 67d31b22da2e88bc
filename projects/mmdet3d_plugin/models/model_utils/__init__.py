from .depthnet import ASPP, DepthNet, HeightNet, Mlp, SELayer

__all__ = ['DepthNet', 'HeightNet', 'ASPP', 'Mlp', 'SELayer']
