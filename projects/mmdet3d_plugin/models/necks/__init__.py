from .fpn import CustomFPN
from .identity import Identity
from .lss_fpn import FPN_LSS
from .lss_heightmap import MGHS, MGHS_Depth, MGHS_Stereo
from .mix import SFA

__all__ = ['SFA', 'Identity', 'MGHS', 'MGHS_Depth', 'MGHS_Stereo', 'FPN_LSS']
