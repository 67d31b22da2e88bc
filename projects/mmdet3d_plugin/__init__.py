"""DHD plugin package: same import root and registry names as the reference's
projects/mmdet3d_plugin (plugin_dir in projects/configs/DHD/DHD-*.py), hot path only."""
from .models import *  # noqa: F401,F403
from .ops import bev_pool_v2  # noqa: F401
