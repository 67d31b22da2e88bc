from .occ_head import predictor

__all__ = ['predictor']
