from .DHD_model import DHD

__all__ = ['DHD']
