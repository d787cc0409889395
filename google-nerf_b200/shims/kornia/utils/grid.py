"""`from kornia.utils.grid import create_meshgrid3d` (ngp_pl/train.py:18)."""
from .. import create_meshgrid, create_meshgrid3d  # noqa: F401
