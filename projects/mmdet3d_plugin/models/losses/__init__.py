from .cross_entropy_loss import CrossEntropyLoss
from .semkitti_loss import geo_scal_loss_with_mask, sem_scal_loss_with_mask

__all__ = ['CrossEntropyLoss', 'sem_scal_loss_with_mask', 'geo_scal_loss_with_mask']
