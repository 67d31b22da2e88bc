from .identity import Identity
from .lss_heightmap import MGHS
from .mix import SFA

__all__ = ['SFA', 'Identity', 'MGHS']
