"""`import tinycudann as tcnn` -> google_nerf_b200.tinycudann (ngp_pl/models/networks.py:4)."""
from google_nerf_b200.tinycudann import Encoding, Module, Network, NetworkWithInputEncoding  # noqa: F401
