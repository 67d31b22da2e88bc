from .bev_pool import QuickCumsumCuda, bev_pool_v2

__all__ = ['bev_pool_v2', 'QuickCumsumCuda']
