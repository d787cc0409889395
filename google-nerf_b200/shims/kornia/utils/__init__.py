from .grid import create_meshgrid, create_meshgrid3d  # noqa: F401
