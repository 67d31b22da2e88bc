"""Drop-in for the reference's ops/bev_pool_v2/bev_pool.py (:11-106): same names, argument
order and return layout; the CUDA work goes through libdhd_b200.so (dhd_bev_pool_v2_fwd/bwd)."""
from dhd_b200.pool import QuickCumsumCuda, bev_pool_v2

__all__ = ['bev_pool_v2', 'QuickCumsumCuda']
