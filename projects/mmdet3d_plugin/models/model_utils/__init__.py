from .depthnet import ASPP, HeightNet, Mlp, SELayer

__all__ = ['HeightNet', 'ASPP', 'Mlp', 'SELayer']
