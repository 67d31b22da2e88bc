from .DHD_model import DHD, DHD_stereo

__all__ = ['DHD', 'DHD_stereo']
