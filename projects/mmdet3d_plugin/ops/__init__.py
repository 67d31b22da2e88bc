from .bev_pool_v2 import bev_pool_v2

__all__ = ['bev_pool_v2']
