"""`import vren` -> google_nerf_b200.vren (ngp_pl/models/custom_functions.py:2, rendering.py:5, networks.py:5)."""
from google_nerf_b200.vren import *  # noqa: F401,F403
from google_nerf_b200.vren import (composite_test_fw, composite_train_bw, composite_train_fw, morton3D,  # noqa: F401
                                   morton3D_invert, packbits, ray_aabb_intersect, ray_sphere_intersect,
                                   raymarching_test, raymarching_train)
